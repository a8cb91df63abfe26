// Host-side transcript formats and the Fiat-Shamir seed chain (f2, f3).  No device code here.
//   * SHA-256 (FIPS 180-4) -- the hash of the Fiat-Shamir chain (the reference lists Fiat-Shamir as TODO, README.md:12, so
//     the chain below is this library's definition; labrador_b200/fs.py restates it with hashlib for the parity tests)
//   * BitWriter / BitReader -- the compact wire format: 13 bits per Z_q coefficient, 2 bits per JL entry
//   * gzip size metric of Transcript::size_in_bytes (structs.rs:211-221) through zlib
#pragma once
#include <stdint.h>
#include <string.h>
#include <zlib.h>
#include <vector>

namespace labwire {

// ---------------------------------------------------------------- SHA-256 ----------------------------------------------------------------
struct Sha256 {
    uint32_t h[8];
    uint8_t buf[64];
    uint64_t len = 0;
    size_t fill = 0;
    Sha256() {
        static const uint32_t iv[8] = {0x6a09e667u, 0xbb67ae85u, 0x3c6ef372u, 0xa54ff53au, 0x510e527fu, 0x9b05688cu, 0x1f83d9abu, 0x5be0cd19u};
        memcpy(h, iv, sizeof h);
    }
    static uint32_t rotr(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }
    void block(const uint8_t *p) {
        static const uint32_t k[64] = {
            0x428a2f98u, 0x71374491u, 0xb5c0fbcfu, 0xe9b5dba5u, 0x3956c25bu, 0x59f111f1u, 0x923f82a4u, 0xab1c5ed5u, 0xd807aa98u, 0x12835b01u, 0x243185beu,
            0x550c7dc3u, 0x72be5d74u, 0x80deb1feu, 0x9bdc06a7u, 0xc19bf174u, 0xe49b69c1u, 0xefbe4786u, 0x0fc19dc6u, 0x240ca1ccu, 0x2de92c6fu, 0x4a7484aau,
            0x5cb0a9dcu, 0x76f988dau, 0x983e5152u, 0xa831c66du, 0xb00327c8u, 0xbf597fc7u, 0xc6e00bf3u, 0xd5a79147u, 0x06ca6351u, 0x14292967u, 0x27b70a85u,
            0x2e1b2138u, 0x4d2c6dfcu, 0x53380d13u, 0x650a7354u, 0x766a0abbu, 0x81c2c92eu, 0x92722c85u, 0xa2bfe8a1u, 0xa81a664bu, 0xc24b8b70u, 0xc76c51a3u,
            0xd192e819u, 0xd6990624u, 0xf40e3585u, 0x106aa070u, 0x19a4c116u, 0x1e376c08u, 0x2748774cu, 0x34b0bcb5u, 0x391c0cb3u, 0x4ed8aa4au, 0x5b9cca4fu,
            0x682e6ff3u, 0x748f82eeu, 0x78a5636fu, 0x84c87814u, 0x8cc70208u, 0x90befffau, 0xa4506cebu, 0xbef9a3f7u, 0xc67178f2u};
        uint32_t w[64];
        for (int i = 0; i < 16; i++) w[i] = (uint32_t)p[4 * i] << 24 | (uint32_t)p[4 * i + 1] << 16 | (uint32_t)p[4 * i + 2] << 8 | p[4 * i + 3];
        for (int i = 16; i < 64; i++) {
            uint32_t s0 = rotr(w[i - 15], 7) ^ rotr(w[i - 15], 18) ^ (w[i - 15] >> 3), s1 = rotr(w[i - 2], 17) ^ rotr(w[i - 2], 19) ^ (w[i - 2] >> 10);
            w[i] = w[i - 16] + s0 + w[i - 7] + s1;
        }
        uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
        for (int i = 0; i < 64; i++) {
            uint32_t S1 = rotr(e, 6) ^ rotr(e, 11) ^ rotr(e, 25), ch = (e & f) ^ (~e & g), t1 = hh + S1 + ch + k[i] + w[i];
            uint32_t S0 = rotr(a, 2) ^ rotr(a, 13) ^ rotr(a, 22), mj = (a & b) ^ (a & c) ^ (b & c), t2 = S0 + mj;
            hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
        }
        h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
    }
    void update(const void *data, size_t n) {
        const uint8_t *p = (const uint8_t *)data;
        len += n;
        while (n) {
            size_t take = 64 - fill < n ? 64 - fill : n;
            memcpy(buf + fill, p, take);
            fill += take; p += take; n -= take;
            if (fill == 64) { block(buf); fill = 0; }
        }
    }
    void final(uint8_t out[32]) {
        uint64_t bits = len * 8;
        uint8_t pad = 0x80;
        update(&pad, 1);
        uint8_t z = 0;
        while (fill != 56) update(&z, 1);
        uint8_t lb[8];
        for (int i = 0; i < 8; i++) lb[i] = (uint8_t)(bits >> (56 - 8 * i));
        update(lb, 8);
        for (int i = 0; i < 8; i++) { out[4 * i] = (uint8_t)(h[i] >> 24); out[4 * i + 1] = (uint8_t)(h[i] >> 16); out[4 * i + 2] = (uint8_t)(h[i] >> 8); out[4 * i + 3] = (uint8_t)h[i]; }
    }
};
inline void sha256(const void *a, size_t na, const void *b, size_t nb, uint8_t out[32]) {
    Sha256 s;
    if (na) s.update(a, na);
    if (nb) s.update(b, nb);
    s.final(out);
}

// --------------------------------------------------------- bit stream (LSB first) ---------------------------------------------------------
struct BitWriter {
    uint8_t *out;
    size_t cap, pos = 0;        // pos in bytes of fully written output
    uint64_t acc = 0;
    int nbits = 0;
    void put(uint32_t v, int bits) {
        acc |= (uint64_t)v << nbits;
        nbits += bits;
        while (nbits >= 8) {
            if (out && pos < cap) out[pos] = (uint8_t)acc;
            pos++; acc >>= 8; nbits -= 8;
        }
    }
    void align() { if (nbits) put(0, 8 - nbits); }
    void raw(const void *p, size_t n) {
        align();
        if (out && pos + n <= cap) memcpy(out + pos, p, n);
        pos += n;
    }
};
struct BitReader {
    const uint8_t *in;
    size_t size, pos = 0;
    uint64_t acc = 0;
    int nbits = 0;
    bool bad = false;
    uint32_t get(int bits) {
        while (nbits < bits) {
            if (pos >= size) { bad = true; return 0; }
            acc |= (uint64_t)in[pos++] << nbits;
            nbits += 8;
        }
        uint32_t v = (uint32_t)(acc & ((1ull << bits) - 1));
        acc >>= bits; nbits -= bits;
        return v;
    }
    void align() { acc = 0; nbits = 0; }
    void raw(void *p, size_t n) {
        align();
        if (pos + n > size) { bad = true; return; }
        memcpy(p, in + pos, n);
        pos += n;
    }
};

// gzip container, best compression (flate2 `GzEncoder::new(_, Compression::best())`, structs.rs:214)
inline int gzip_size(const uint8_t *data, size_t n, size_t *out_size) {
    z_stream zs;
    memset(&zs, 0, sizeof zs);
    if (deflateInit2(&zs, 9, Z_DEFLATED, 15 + 16, 8, Z_DEFAULT_STRATEGY) != Z_OK) return -1;
    std::vector<uint8_t> buf(1 << 16);
    size_t total = 0, off = 0;
    int rc = Z_OK;
    do {
        const size_t take = n - off < ((size_t)1 << 30) ? n - off : ((size_t)1 << 30);
        zs.next_in = const_cast<Bytef *>(data + off);
        zs.avail_in = (uInt)take;
        off += take;
        const int flush = off == n ? Z_FINISH : Z_NO_FLUSH;
        do {
            zs.next_out = buf.data();
            zs.avail_out = (uInt)buf.size();
            rc = deflate(&zs, flush);
            if (rc == Z_STREAM_ERROR) { deflateEnd(&zs); return -1; }
            total += buf.size() - zs.avail_out;
        } while (zs.avail_out == 0);
    } while (rc != Z_STREAM_END);
    deflateEnd(&zs);
    *out_size = total;
    return 0;
}

}  // namespace labwire
