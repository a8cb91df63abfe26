// Exercises labrador-snark_b200/cpp/labrador.hpp (the C++ host mirror of the reference's prover API) against values the
// pytest driver computed with the CPU oracle.  Usage: test_labrador_hpp <case.bin>   (tests/test_gpu_parity.py writes the file)
// With no argument it only checks what needs no GPU (constants, packing, that the library refuses to run without a device).
#include <cstdio>
#include <cstring>
#include <fstream>
#include "../../labrador-snark_b200/cpp/labrador.hpp"

using namespace labrador;

template <typename T>
static void rd(std::ifstream &f, std::vector<T> &v, std::size_t n) { v.resize(n); f.read(reinterpret_cast<char *>(v.data()), n * sizeof(T)); }
template <typename T>
static void rd(std::ifstream &f, T *p, std::size_t n) { f.read(reinterpret_cast<char *>(p), n * sizeof(T)); }
static int fail(const char *what) { std::printf("MISMATCH %s\n", what); return 1; }

int main(int argc, char **argv) {
    RuntimeConstants c = RuntimeConstants::make(2, 2);
    if (c.KAPPA != 128 || c.T_1 != 4 || c.B_1 != 9 || c.T_2 != 2 || c.B_2 != 14 || c.BETA_BOUND != 31) return fail("RuntimeConstants::new(2,2)");
    try { RuntimeConstants::make(4096, 64); return fail("degenerate constants accepted"); } catch (const Error &) {}
    Challenges pk;
    pk.pi = {1, -1, 0, 0, 1, 1, -1, -1, 0, 0, 0, 0, 0, 0, 0, 1};
    pk.pack();
    if (pk.pi2.size() != 1 || pk.pi2[0] != (0x8031u | (0x00C2u << 16))) return fail("lab_pi_pack");
    if (argc < 2) { std::printf("OK (host-only checks)\n"); return 0; }

    std::ifstream f(argv[1], std::ios::binary);
    if (!f) return fail("cannot open the case file");
    std::uint64_t hdr[4];                                   // N, R, n_attempts, expected jl_attempt
    rd(f, hdr, 4);
    const std::size_t N = hdr[0], R = hdr[1], K = N * D;
    c = RuntimeConstants::make(N, R);
    std::array<std::uint8_t, 32> seed{};
    rd(f, seed.data(), 32);
    std::vector<std::uint32_t> S;
    State st;
    Challenges ch;
    rd(f, S, R * N * D); rd(f, st.phi, R * N * D); rd(f, st.a, R * R * D); rd(f, st.b.data(), D);
    ch.n_attempts = (int)hdr[2];
    rd(f, ch.pi, hdr[2] * R * LAB_JL_ROWS * N * D);
    rd(f, &ch.psi, 1); rd(f, ch.omega.data(), LAB_JL_ROWS); rd(f, ch.alpha.data(), D); rd(f, ch.beta.data(), D); rd(f, ch.c, R * D);
    Transcript want;
    rd(f, want.u_1, K * D); rd(f, want.u_2, K * D); rd(f, want.z, N * D); rd(f, want.t_i_all, R * K * D); rd(f, want.g_mat, R * R * D); rd(f, want.h_mat, R * R * D);
    rd(f, want.projection_int.data(), LAB_JL_ROWS); rd(f, want.b_prime_prime.data(), D);
    std::uint64_t bincode_len = 0;
    rd(f, &bincode_len, 1);
    if (!f) return fail("short case file");

    Context ctx(0);
    CRS crs = CRS::from_seed(seed, c);
    Prover prover(S, c);
    Transcript tr = prover.proof_gen(ctx, st, crs, ch);
    if (tr.jl_attempt != (int)hdr[3]) return fail("jl_attempt");
    if (tr.u_1 != want.u_1) return fail("u_1");
    if (tr.u_2 != want.u_2) return fail("u_2");
    if (tr.z != want.z) return fail("z");
    if (tr.t_i_all != want.t_i_all) return fail("t_i_all");
    if (tr.g_mat != want.g_mat) return fail("g_mat");
    if (tr.h_mat != want.h_mat) return fail("h_mat");
    if (tr.projection_int != want.projection_int) return fail("projection");
    if (tr.b_prime_prime != want.b_prime_prime) return fail("b_prime_prime");
    if (prover.verify(ctx, st, crs, ch, tr) != 0) return fail("verify rejected an honest transcript");
    if (prover.to_bincode(ch, tr).size() != bincode_len) return fail("bincode length");
    if (prover.size_in_bytes(ch, tr) >= bincode_len) return fail("gzip size");
    Challenges ch2 = ch;
    ch2.pack();                                             // the same proof with the JL attempts 2-bit packed
    Transcript tr2 = prover.proof_gen(ctx, st, crs, ch2);
    if (tr2.u_1 != want.u_1 || tr2.z != want.z || tr2.h_mat != want.h_mat) return fail("packed challenges");
    tr.u_2[5] ^= 1u;
    if (prover.verify(ctx, st, crs, ch, tr) != 20) return fail("tampered u_2 must fail check 20");
    auto row = crs.fetch_A_row(ctx, 3);
    auto prod = polymul_batch(ctx, row, row);
    if (prod.size() != N) return fail("polymul_batch");
    std::printf("OK\n");
    return 0;
}
