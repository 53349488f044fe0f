// kgv_blake2b.cuh — keyed BLAKE2b-256 streaming hasher, one hash per thread.
//
// GPU counterpart of the reference's domain hashers (crypto/hashes/src/hashers.rs:23-33,77-106:
// blake2b_simd Params::new().hash_length(32).key(domain)) for the three domains on the path:
// TransactionID, TransactionHash, TransactionSigningHash; plus unkeyed BLAKE2b-256 for the P2SH
// script hash (crypto/txscript/src/standard.rs:50-54, opcodes/mod.rs:738-744).
// The state after the (constant) key block is precomputed per domain (tools/derive_constants.py),
// so a keyed hash costs one compression less than a literal implementation.
#pragma once
#include "kgv_arith.cuh"

namespace kgv {

enum Blake2bDomain { B2B_TX_ID = 0, B2B_TX_HASH = 1, B2B_SIGHASH = 2, B2B_UNKEYED = 3 };

#if defined(__CUDACC__)
#define KGV_CONST_TABLE static __device__ __constant__
#else
#define KGV_CONST_TABLE static const
#endif

KGV_CONST_TABLE uint64_t kB2bIV[8] = {0x6a09e667f3bcc908ull, 0xbb67ae8584caa73bull, 0x3c6ef372fe94f82bull, 0xa54ff53a5f1d36f1ull,
                                      0x510e527fade682d1ull, 0x9b05688c2b3e6c1full, 0x1f83d9abfb41bd6bull, 0x5be0cd19137e2179ull};
// chaining value after compressing the zero-padded key block (t = 128, not last)
KGV_CONST_TABLE uint64_t kB2bMid[3][8] = {
    {0x2068b145a3adacdeull, 0x6e1b0076db22d364ull, 0x9696096cf8961ec3ull, 0x595c2e7dd903d7d6ull, 0xa4f2bbb08a23797aull, 0x90a8690c83736248ull,
     0x41dcb76629a54308ull, 0xefc33e65686ca8f6ull},
    {0x74daa94e84e5e80dull, 0x72a5dc8b653aaebbull, 0xe6114975b79cd6f4ull, 0xd53c4b5f24409893ull, 0x8aac4213dc4a4f71ull, 0x451d53793201eadfull,
     0xb39c1487921e0d06ull, 0x702dc6a2a02b2701ull},
    {0xa1f262d8452f7944ull, 0xa0c8b06e8fc22fe0ull, 0xdfc64e0f0abc65e6ull, 0xfc7f6ae21f29953full, 0x8e5ddbf833addef9ull, 0xb0fdcb71ba2162beull,
     0x9215ad9b69ac9b9aull, 0x0218557f8b2f7d32ull}};
// digest of the empty message per keyed domain (the key block is then the last block)
KGV_CONST_TABLE uint64_t kB2bEmpty[3][4] = {
    {0x2b4d89a0fd5ef6e5ull, 0xcc9a6ee4e9c29005ull, 0xf522155a502f0303ull, 0x9c1d0bc75e8cc7e8ull},
    {0x0228c7379e2a2750ull, 0x6744aba6edd0936full, 0x487690878b3327f6ull, 0xbfb33c1991d2883cull},
    {0x0d7462ad3750c734ull, 0x794f848ff828324bull, 0x45a055cdfa7bc001ull, 0xce525ec1ab8e51beull}};
KGV_CONST_TABLE uint8_t kB2bSigma[12][16] = {
    {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3},
    {11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4}, {7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8},
    {9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13}, {2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9},
    {12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11}, {13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10},
    {6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5}, {10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0},
    {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3}};

KGV_HD uint64_t rotr64(uint64_t x, int r) { return (x >> r) | (x << (64 - r)); }

struct Blake2b {
  uint64_t h[8];
  uint64_t m[16];   // current (partially filled) block, little-endian words
  uint64_t t;       // bytes compressed so far + bytes in m
  uint32_t fill;    // bytes in m
  uint32_t domain;  // Blake2bDomain
  bool fresh;       // keyed and nothing absorbed yet
};

#if defined(__CUDACC__)
static __device__ __noinline__
#else
inline
#endif
void b2b_compress(uint64_t* h, const uint64_t* m, uint64_t t, bool last) {
  uint64_t v[16];
#pragma unroll
  for (int i = 0; i < 8; i++) { v[i] = h[i]; v[8 + i] = kB2bIV[i]; }
  v[12] ^= t;
  if (last) v[14] = ~v[14];
#define KGV_B2B_G(a, b, c, d, x, y)     \
  v[a] = v[a] + v[b] + (x); v[d] = rotr64(v[d] ^ v[a], 32); \
  v[c] = v[c] + v[d];       v[b] = rotr64(v[b] ^ v[c], 24); \
  v[a] = v[a] + v[b] + (y); v[d] = rotr64(v[d] ^ v[a], 16); \
  v[c] = v[c] + v[d];       v[b] = rotr64(v[b] ^ v[c], 63);
  for (int r = 0; r < 12; r++) {
    const uint8_t* s = kB2bSigma[r];
    KGV_B2B_G(0, 4, 8, 12, m[s[0]], m[s[1]]);
    KGV_B2B_G(1, 5, 9, 13, m[s[2]], m[s[3]]);
    KGV_B2B_G(2, 6, 10, 14, m[s[4]], m[s[5]]);
    KGV_B2B_G(3, 7, 11, 15, m[s[6]], m[s[7]]);
    KGV_B2B_G(0, 5, 10, 15, m[s[8]], m[s[9]]);
    KGV_B2B_G(1, 6, 11, 12, m[s[10]], m[s[11]]);
    KGV_B2B_G(2, 7, 8, 13, m[s[12]], m[s[13]]);
    KGV_B2B_G(3, 4, 9, 14, m[s[14]], m[s[15]]);
  }
#undef KGV_B2B_G
#pragma unroll
  for (int i = 0; i < 8; i++) h[i] ^= v[i] ^ v[8 + i];
}

KGV_HD void b2b_init(Blake2b& s, uint32_t domain) {
  s.domain = domain;
  s.fill = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) s.m[i] = 0;
  if (domain == B2B_UNKEYED) {
#pragma unroll
    for (int i = 0; i < 8; i++) s.h[i] = kB2bIV[i];
    s.h[0] ^= 0x01010020ull;  // digest 32, no key, fanout 1, depth 1
    s.t = 0;
    s.fresh = false;
  } else {
#pragma unroll
    for (int i = 0; i < 8; i++) s.h[i] = kB2bMid[domain][i];
    s.t = 128;
    s.fresh = true;
  }
}

KGV_HD void b2b_byte(Blake2b& s, uint32_t byte) {
  if (s.fill == 128) {  // buffer full and more input follows: not the last block
    b2b_compress(s.h, s.m, s.t, false);
    s.fill = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s.m[i] = 0;
  }
  s.m[s.fill >> 3] |= (uint64_t)(byte & 0xFFu) << (8 * (s.fill & 7));
  s.fill++;
  s.t++;
  s.fresh = false;
}
KGV_HD void b2b_u8(Blake2b& s, uint32_t v) { b2b_byte(s, v); }
KGV_HD void b2b_u16(Blake2b& s, uint32_t v) { b2b_byte(s, v); b2b_byte(s, v >> 8); }
KGV_HD void b2b_u32(Blake2b& s, uint32_t v) {
#pragma unroll
  for (int i = 0; i < 4; i++) b2b_byte(s, v >> (8 * i));
}
KGV_HD void b2b_u64(Blake2b& s, uint64_t v) {
  if ((s.fill & 7) == 0 && s.fill < 128) {  // aligned fast path
    s.m[s.fill >> 3] = v;
    s.fill += 8;
    s.t += 8;
    s.fresh = false;
    return;
  }
#pragma unroll
  for (int i = 0; i < 8; i++) b2b_byte(s, (uint32_t)(v >> (8 * i)));
}
KGV_HD void b2b_bytes(Blake2b& s, const uint8_t* p, uint32_t n) {
  for (uint32_t i = 0; i < n; i++) b2b_byte(s, p[i]);
}
// 32 bytes given as 4 little-endian u64 words (digests produced by b2b_final)
KGV_HD void b2b_digest_words(Blake2b& s, const uint64_t* w) {
#pragma unroll
  for (int i = 0; i < 4; i++) b2b_u64(s, w[i]);
}
// out: 4 little-endian u64 words = the 32 digest bytes
KGV_HD void b2b_final(Blake2b& s, uint64_t* out) {
  if (s.fresh) {  // keyed hash of the empty message
#pragma unroll
    for (int i = 0; i < 4; i++) out[i] = kB2bEmpty[s.domain][i];
    return;
  }
  b2b_compress(s.h, s.m, s.t, true);
#pragma unroll
  for (int i = 0; i < 4; i++) out[i] = s.h[i];
}

}  // namespace kgv
