// kgv_sha256.cuh — SHA-256 compression for the BIP-340 challenge hash and the ECDSA sighash wrap.
// Reference sites: BIP-340 tagged hash inside libsecp256k1's schnorrsig_verify (reached from
// crypto/txscript/src/lib.rs:593) and TransactionSigningHashECDSA (crypto/hashes/src/hashers.rs:35-75).
#pragma once
#include "kgv_arith.cuh"

namespace kgv {

#if defined(__CUDACC__)
static __device__ __constant__ uint32_t kSha256K[64] = {
#else
static const uint32_t kSha256K[64] = {
#endif
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01, 0x243185be,
    0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa,
    0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967, 0x27b70a85,
    0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85, 0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3,
    0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f,
    0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};

KGV_HD uint32_t rotr32(uint32_t x, int r) {
#if defined(__CUDACC__)
  return __funnelshift_r(x, x, r);
#else
  return (x >> r) | (x << (32 - r));
#endif
}

// one compression; w = 16 big-endian message words (destroyed: used as the rolling schedule)
KGV_HD void sha256_compress(uint32_t* st, uint32_t* w) {
  uint32_t a = st[0], b = st[1], c = st[2], d = st[3], e = st[4], f = st[5], g = st[6], h = st[7];
#pragma unroll
  for (int i = 0; i < 64; i++) {
    uint32_t wi;
    if (i < 16) {
      wi = w[i];
    } else {
      uint32_t w15 = w[(i + 1) & 15], w2 = w[(i + 14) & 15];
      uint32_t s0 = rotr32(w15, 7) ^ rotr32(w15, 18) ^ (w15 >> 3);
      uint32_t s1 = rotr32(w2, 17) ^ rotr32(w2, 19) ^ (w2 >> 10);
      wi = w[i & 15] + s0 + w[(i + 9) & 15] + s1;
      w[i & 15] = wi;
    }
    uint32_t S1 = rotr32(e, 6) ^ rotr32(e, 11) ^ rotr32(e, 25);
    uint32_t ch = (e & f) ^ (~e & g);
    uint32_t t1 = h + S1 + ch + kSha256K[i] + wi;
    uint32_t S0 = rotr32(a, 2) ^ rotr32(a, 13) ^ rotr32(a, 22);
    uint32_t mj = (a & b) ^ (a & c) ^ (b & c);
    uint32_t t2 = S0 + mj;
    h = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
  }
  st[0] += a; st[1] += b; st[2] += c; st[3] += d; st[4] += e; st[5] += f; st[6] += g; st[7] += h;
}

// e = SHA256(SHA256(tag)||SHA256(tag)||r||pk||m) with tag = "BIP0340/challenge".
// rw, pkw, mw: 8 big-endian words each.  out: 8 big-endian words.
KGV_HD void bip340_challenge(uint32_t* out, const uint32_t* rw, const uint32_t* pkw, const uint32_t* mw) {
  // state after the 64-byte tag block (tools/derive_constants.py)
  uint32_t st[8] = {0x9cecba11u, 0x23925381u, 0x11679112u, 0xd1627e0fu, 0x97c87550u, 0x003cc765u, 0x90f61164u, 0x33e9b66au};
  uint32_t w[16];
#pragma unroll
  for (int i = 0; i < 8; i++) { w[i] = rw[i]; w[8 + i] = pkw[i]; }
  sha256_compress(st, w);
#pragma unroll
  for (int i = 0; i < 8; i++) w[i] = mw[i];
  w[8] = 0x80000000u;
#pragma unroll
  for (int i = 9; i < 15; i++) w[i] = 0;
  w[15] = 160 * 8;  // 64-byte tag block + 96 bytes
  sha256_compress(st, w);
#pragma unroll
  for (int i = 0; i < 8; i++) out[i] = st[i];
}

// SHA256( SHA256("TransactionSigningHashECDSA") || h )  (hashers.rs:39-60, sighash.rs:267-277)
KGV_HD void ecdsa_sighash_wrap(uint32_t* out, const uint32_t* hw) {
  uint32_t st[8] = {0x6a09e667u, 0xbb67ae85u, 0x3c6ef372u, 0xa54ff53au, 0x510e527fu, 0x9b05688cu, 0x1f83d9abu, 0x5be0cd19u};
  uint32_t w[16] = {0xa4f2ece4u, 0x5a286cb1u, 0xec0a4e4du, 0x383468d0u, 0x00f71757u, 0x052b1504u, 0xaa349532u, 0x8df5f4eau};
#pragma unroll
  for (int i = 0; i < 8; i++) w[8 + i] = hw[i];
  sha256_compress(st, w);
  w[0] = 0x80000000u;
#pragma unroll
  for (int i = 1; i < 15; i++) w[i] = 0;
  w[15] = 64 * 8;
  sha256_compress(st, w);
#pragma unroll
  for (int i = 0; i < 8; i++) out[i] = st[i];
}

}  // namespace kgv
