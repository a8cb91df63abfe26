// labrador.hpp -- C++17 host mirror of the reference's prover API over the C ABI (header only).
// Same names and argument meaning as the reference (constants.rs / structs.rs / proofgen.rs); errors that are
// panics in the reference are exceptions here.  Dense layouts as documented in include/labrador_b200.h.
#pragma once
#include <array>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>
#include "../../include/labrador_b200.h"

namespace labrador {

constexpr std::size_t D = LAB_D;
constexpr std::uint32_t Q = LAB_Q;

struct Error : std::runtime_error {
    int status;
    Error(int s, const std::string &m) : std::runtime_error(m), status(s) {}
};

class Context {
  public:
    explicit Context(int device = 0) {
        if (int rc = lab_ctx_create(device, &ctx_); rc != LAB_OK) throw Error(rc, lab_last_error(nullptr));
    }
    ~Context() { lab_ctx_destroy(ctx_); }
    Context(const Context &) = delete;
    Context &operator=(const Context &) = delete;
    lab_ctx *get() const { return ctx_; }
    void check(int rc) const { if (rc != LAB_OK) throw Error(rc, lab_last_error(ctx_)); }
    // one process per GPU: rank 0 makes the id, the host distributes it, every rank attaches (collective); proof_gen /
    // verify then shard the CRS-regenerating stages by rows inside the library
    static std::array<std::uint8_t, LAB_COMM_ID_BYTES> comm_unique_id() {
        std::array<std::uint8_t, LAB_COMM_ID_BYTES> id{};
        if (int rc = lab_comm_unique_id(id.data()); rc != LAB_OK) throw Error(rc, lab_last_error(nullptr));
        return id;
    }
    void comm_init(const std::array<std::uint8_t, LAB_COMM_ID_BYTES> &id, int rank, int world) const { check(lab_comm_init(ctx_, id.data(), rank, world)); }
    // keep transformed CRS polynomials in HBM between calls (verify after prove, proofs under one CRS); 0 = off
    void crs_cache_configure(std::size_t max_bytes) const { check(lab_crs_cache_configure(ctx_, max_bytes)); }
  private:
    lab_ctx *ctx_ = nullptr;
};

// RuntimeConstants::new(N, R)  (constants.rs:234-264)
struct RuntimeConstants : lab_constants {
    static RuntimeConstants make(std::uint64_t N, std::uint64_t R) {
        RuntimeConstants c{};
        if (int rc = lab_runtime_constants(N, R, &c); rc != LAB_OK) throw Error(rc, "degenerate RuntimeConstants (SURVEY F8)");
        return c;
    }
};

using Poly = std::array<std::uint32_t, D>;          // dense Rq

// CRS (structs.rs:27-190); from_seed is the parity addition
struct CRS {
    std::array<std::uint8_t, 32> base_seed{};
    const RuntimeConstants *constants = nullptr;
    static CRS from_seed(const std::array<std::uint8_t, 32> &seed, const RuntimeConstants &c) { return CRS{seed, &c}; }
    std::vector<Poly> fetch(const Context &ctx, int which, std::uint64_t i, std::uint64_t j, std::uint64_t k, std::uint64_t row) const {
        std::size_t n = which == 'A' ? constants->N : (which == 'B' ? constants->KAPPA : constants->KAPPA_2);
        std::vector<Poly> out(n);
        ctx.check(lab_crs_fetch(ctx.get(), constants, base_seed.data(), which, i, j, k, row, out[0].data()));
        return out;
    }
    std::vector<Poly> fetch_A_row(const Context &ctx, std::uint64_t row) const { return fetch(ctx, 'A', 0, 0, 0, row); }
    std::vector<Poly> fetch_B_ik_row(const Context &ctx, std::uint64_t i, std::uint64_t k, std::uint64_t row) const { return fetch(ctx, 'B', i, 0, k, row); }
    std::vector<Poly> fetch_C_ijk(const Context &ctx, std::uint64_t i, std::uint64_t j, std::uint64_t k) const { return fetch(ctx, 'C', i, j, k, 0); }
    std::vector<Poly> fetch_D_ijk(const Context &ctx, std::uint64_t i, std::uint64_t j, std::uint64_t k) const { return fetch(ctx, 'D', i, j, k, 0); }
};

// State (structs.rs:269-286), K = L = 1
struct State {
    std::vector<std::uint32_t> phi;   // [R][N][64]
    std::vector<std::uint32_t> a;     // [R][R][64]
    Poly b{};
};

// verifier randomness in consumption order
struct Challenges {
    std::vector<std::int8_t> pi;      // [n_attempts][R][256][N*64]
    std::vector<std::uint32_t> pi2;   // the same 2-bit packed, [n_attempts][R][256][N*4]; used instead of pi when not empty
    int n_attempts = 1;
    std::uint32_t psi = 0;
    std::array<std::uint32_t, LAB_JL_ROWS> omega{};
    Poly alpha{}, beta{};
    std::vector<std::uint32_t> c;     // [R][64]
    lab_challenges raw() const {
        return lab_challenges{pi2.empty() ? pi.data() : nullptr, n_attempts, psi, omega.data(), alpha.data(), beta.data(), c.data(), pi2.empty() ? nullptr : pi2.data()};
    }
    // int8 entries {-1,0,1} -> packed words (lab_pi_pack)
    void pack() {
        pi2.resize(pi.size() / 16);
        if (int rc = lab_pi_pack(pi.data(), pi.size(), pi2.data()); rc != LAB_OK) throw Error(rc, "lab_pi_pack: entries must be in {-1,0,1}");
    }
};

// Transcript (structs.rs:192-209)
struct Transcript {
    std::vector<std::uint32_t> u_1, u_2, z, t_i_all, g_mat, h_mat, phi_final;
    std::array<std::int64_t, LAB_JL_ROWS> projection_int{};
    std::array<std::uint32_t, LAB_JL_ROWS> projection{};
    Poly b_prime_prime{};
    int jl_attempt = 0;
    std::uint64_t norm_sum = 0;
};

// Prover (proofgen.rs:14-28)
class Prover {
  public:
    Prover(const std::vector<std::uint32_t> &witness /* [R][N][64] */, const RuntimeConstants &c) : witness_(witness), c_(c) {}

    // Prover::proof_gen (proofgen.rs:30-427)
    Transcript proof_gen(const Context &ctx, const State &st, const CRS &crs, const Challenges &ch) const {
        const std::size_t R = c_.R, N = c_.N, K = c_.KAPPA;
        Transcript tr;
        tr.u_1.resize(K * D); tr.u_2.resize(K * D); tr.z.resize(N * D); tr.t_i_all.resize(R * K * D);
        tr.g_mat.resize(R * R * D); tr.h_mat.resize(R * R * D); tr.phi_final.resize(R * N * D);
        lab_state cst{st.phi.data(), st.a.data(), st.b.data()};
        lab_challenges cch = ch.raw();
        lab_transcript out{tr.u_1.data(), 0, tr.projection_int.data(), tr.projection.data(), tr.b_prime_prime.data(), tr.u_2.data(),
                           tr.z.data(), tr.t_i_all.data(), tr.g_mat.data(), tr.h_mat.data(), tr.phi_final.data(), 0};
        int rc = lab_prove(ctx.get(), &c_, crs.base_seed.data(), witness_.data(), &cst, &cch, &out);
        if (rc == LAB_ERR_JL_REJECTED) throw Error(rc, "failed JL...");                       // proofgen.rs:176
        ctx.check(rc);
        tr.jl_attempt = out.jl_attempt;
        tr.norm_sum = out.norm_sum;
        return tr;
    }

    // Verifier::verify (verification.rs:25-438): returns the reference's check number that failed, 0 = accepted
    int verify(const Context &ctx, const State &st, const CRS &crs, const Challenges &ch, Transcript &tr) const {
        lab_state cst{st.phi.data(), st.a.data(), st.b.data()};
        lab_challenges cch = ch.raw();
        lab_transcript in{tr.u_1.data(), tr.jl_attempt, tr.projection_int.data(), tr.projection.data(), tr.b_prime_prime.data(), tr.u_2.data(),
                          tr.z.data(), tr.t_i_all.data(), tr.g_mat.data(), tr.h_mat.data(), nullptr, 0};
        int accepted = 0, failed = 0;
        std::uint64_t norm = 0;
        ctx.check(lab_verify(ctx.get(), &c_, crs.base_seed.data(), &cst, &cch, &in, &accepted, &failed, &norm));
        return accepted ? 0 : failed;
    }
    // Transcript::size_in_bytes (structs.rs:211-221): gzip(best) of the bincode bytes
    std::size_t size_in_bytes(const Challenges &ch, Transcript &tr) const {
        lab_challenges cch = ch.raw();
        lab_transcript in{tr.u_1.data(), tr.jl_attempt, tr.projection_int.data(), tr.projection.data(), tr.b_prime_prime.data(), tr.u_2.data(),
                          tr.z.data(), tr.t_i_all.data(), tr.g_mat.data(), tr.h_mat.data(), nullptr, 0};
        std::size_t gz = 0, raw = 0;
        if (int rc = lab_transcript_size_in_bytes(&c_, &in, &cch, &gz, &raw); rc != LAB_OK) throw Error(rc, "lab_transcript_size_in_bytes");
        return gz;
    }
    // bincode::serialize(&Transcript) (structs.rs:192-221)
    std::vector<std::uint8_t> to_bincode(const Challenges &ch, Transcript &tr) const {
        lab_challenges cch = ch.raw();
        lab_transcript in{tr.u_1.data(), tr.jl_attempt, tr.projection_int.data(), tr.projection.data(), tr.b_prime_prime.data(), tr.u_2.data(),
                          tr.z.data(), tr.t_i_all.data(), tr.g_mat.data(), tr.h_mat.data(), nullptr, 0};
        std::size_t size = 0;
        if (int rc = lab_transcript_bincode(&c_, &in, &cch, nullptr, 0, &size); rc != LAB_OK) throw Error(rc, "lab_transcript_bincode");
        std::vector<std::uint8_t> out(size);
        if (int rc = lab_transcript_bincode(&c_, &in, &cch, out.data(), out.size(), &size); rc != LAB_OK) throw Error(rc, "lab_transcript_bincode");
        return out;
    }

  private:
    const std::vector<std::uint32_t> &witness_;
    const RuntimeConstants &c_;
};

// &Rq * &Rq for a batch (algebraic.rs:517-523), polynomial_vec_inner_product (util.rs:496-509)
inline std::vector<Poly> polymul_batch(const Context &ctx, const std::vector<Poly> &a, const std::vector<Poly> &b) {
    if (a.size() != b.size()) throw Error(LAB_ERR_SHAPE, "polymul_batch: unequal lengths");
    std::vector<Poly> c(a.size());
    if (!a.empty()) ctx.check(lab_polymul_batch(ctx.get(), a[0].data(), b[0].data(), c[0].data(), a.size()));
    return c;
}
inline Poly polynomial_vec_inner_product(const Context &ctx, const std::vector<Poly> &v1, const std::vector<Poly> &v2) {
    if (v1.size() != v2.size()) throw Error(LAB_ERR_SHAPE, "inner product not defined on vectors of unequal length");
    Poly out{};
    if (!v1.empty()) ctx.check(lab_inner_product_batch(ctx.get(), v1[0].data(), v2[0].data(), 1, v1.size(), out.data()));
    return out;
}

}  // namespace labrador
