/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the shipped product path.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library.
 *
 * CPU restatement of the hash primitives the reference's validation path uses:
 *   - SHA-256 (FIPS 180-4)            -> reference uses the `sha2 0.10.8` crate
 *     (crypto/hashes/src/hashers.rs:35-75: sha256 hashers with a pre-hashed domain block)
 *   - BLAKE2b (RFC 7693), keyed, 32-byte digest -> reference uses `blake2b_simd 1.0.2`
 *     (crypto/hashes/src/hashers.rs:77-106: Params::new().hash_length(32).key(domain))
 * Neither crate is vendored under /root/reference; the algorithms are restated from their
 * published specifications and pinned by the reference's own vectors
 * (crypto/hashes/src/hashers.rs:142-233), see tests/test_oracle_golden.py.
 */
#include "ok_oracle.h"
#include <string.h>

/* ---------------------------------------------------------------- SHA-256 */
static const uint32_t K256[64] = {
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01, 0x243185be,
    0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa,
    0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967, 0x27b70a85,
    0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85, 0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3,
    0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f,
    0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};

static inline uint32_t ror32(uint32_t x, int r) { return (x >> r) | (x << (32 - r)); }

static void sha256_compress(uint32_t st[8], const uint8_t blk[64]) {
  uint32_t w[64];
  for (int i = 0; i < 16; i++) w[i] = ((uint32_t)blk[4 * i] << 24) | ((uint32_t)blk[4 * i + 1] << 16) | ((uint32_t)blk[4 * i + 2] << 8) | blk[4 * i + 3];
  for (int i = 16; i < 64; i++) {
    uint32_t s0 = ror32(w[i - 15], 7) ^ ror32(w[i - 15], 18) ^ (w[i - 15] >> 3);
    uint32_t s1 = ror32(w[i - 2], 17) ^ ror32(w[i - 2], 19) ^ (w[i - 2] >> 10);
    w[i] = w[i - 16] + s0 + w[i - 7] + s1;
  }
  uint32_t a = st[0], b = st[1], c = st[2], d = st[3], e = st[4], f = st[5], g = st[6], h = st[7];
  for (int i = 0; i < 64; i++) {
    uint32_t S1 = ror32(e, 6) ^ ror32(e, 11) ^ ror32(e, 25);
    uint32_t ch = (e & f) ^ (~e & g);
    uint32_t t1 = h + S1 + ch + K256[i] + w[i];
    uint32_t S0 = ror32(a, 2) ^ ror32(a, 13) ^ ror32(a, 22);
    uint32_t mj = (a & b) ^ (a & c) ^ (b & c);
    uint32_t t2 = S0 + mj;
    h = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
  }
  st[0] += a; st[1] += b; st[2] += c; st[3] += d; st[4] += e; st[5] += f; st[6] += g; st[7] += h;
}

void ok_sha256_init(ok_sha256_ctx* c) {
  static const uint32_t iv[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
  memcpy(c->st, iv, sizeof iv);
  c->len = 0;
}

void ok_sha256_update(ok_sha256_ctx* c, const void* data, size_t n) {
  const uint8_t* p = (const uint8_t*)data;
  size_t fill = (size_t)(c->len & 63);
  c->len += n;
  if (fill) {
    size_t take = 64 - fill;
    if (take > n) take = n;
    memcpy(c->buf + fill, p, take);
    p += take; n -= take; fill += take;
    if (fill < 64) return;
    sha256_compress(c->st, c->buf);
  }
  while (n >= 64) { sha256_compress(c->st, p); p += 64; n -= 64; }
  if (n) memcpy(c->buf, p, n);
}

void ok_sha256_final(ok_sha256_ctx* c, uint8_t out[32]) {
  uint64_t bits = c->len * 8;
  size_t fill = (size_t)(c->len & 63);
  c->buf[fill++] = 0x80;
  if (fill > 56) { memset(c->buf + fill, 0, 64 - fill); sha256_compress(c->st, c->buf); fill = 0; }
  memset(c->buf + fill, 0, 56 - fill);
  for (int i = 0; i < 8; i++) c->buf[56 + i] = (uint8_t)(bits >> (56 - 8 * i));
  sha256_compress(c->st, c->buf);
  for (int i = 0; i < 8; i++) { out[4 * i] = (uint8_t)(c->st[i] >> 24); out[4 * i + 1] = (uint8_t)(c->st[i] >> 16); out[4 * i + 2] = (uint8_t)(c->st[i] >> 8); out[4 * i + 3] = (uint8_t)c->st[i]; }
}

void ok_sha256(const void* data, size_t n, uint8_t out[32]) {
  ok_sha256_ctx c;
  ok_sha256_init(&c);
  ok_sha256_update(&c, data, n);
  ok_sha256_final(&c, out);
}

/* Reference: crypto/hashes/src/hashers.rs:39-60 — sha256 hashers start from
 * SHA256(domain) fed as the first 32 bytes of the stream. */
void ok_sha256_domain(const char* domain, const void* data, size_t n, uint8_t out[32]) {
  uint8_t dh[32];
  ok_sha256(domain, strlen(domain), dh);
  ok_sha256_ctx c;
  ok_sha256_init(&c);
  ok_sha256_update(&c, dh, 32);
  ok_sha256_update(&c, data, n);
  ok_sha256_final(&c, out);
}

/* ---------------------------------------------------------------- BLAKE2b */
static const uint64_t B2B_IV[8] = {0x6a09e667f3bcc908ULL, 0xbb67ae8584caa73bULL, 0x3c6ef372fe94f82bULL, 0xa54ff53a5f1d36f1ULL,
                                   0x510e527fade682d1ULL, 0x9b05688c2b3e6c1fULL, 0x1f83d9abfb41bd6bULL, 0x5be0cd19137e2179ULL};
static const uint8_t B2B_SIGMA[12][16] = {
    {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3},
    {11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4}, {7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8},
    {9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13}, {2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9},
    {12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11}, {13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10},
    {6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5}, {10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0},
    {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3}};

static inline uint64_t ror64(uint64_t x, int r) { return (x >> r) | (x << (64 - r)); }

static void b2b_compress(ok_blake2b_ctx* c, const uint8_t blk[128], int last) {
  uint64_t m[16], v[16];
  for (int i = 0; i < 16; i++) {
    uint64_t w = 0;
    for (int j = 7; j >= 0; j--) w = (w << 8) | blk[8 * i + j];
    m[i] = w;
  }
  for (int i = 0; i < 8; i++) { v[i] = c->h[i]; v[i + 8] = B2B_IV[i]; }
  v[12] ^= c->t[0];
  v[13] ^= c->t[1];
  if (last) v[14] = ~v[14];
#define G(a, b, cc, d, x, y)                 \
  v[a] = v[a] + v[b] + (x); v[d] = ror64(v[d] ^ v[a], 32); \
  v[cc] = v[cc] + v[d];     v[b] = ror64(v[b] ^ v[cc], 24); \
  v[a] = v[a] + v[b] + (y); v[d] = ror64(v[d] ^ v[a], 16); \
  v[cc] = v[cc] + v[d];     v[b] = ror64(v[b] ^ v[cc], 63);
  for (int r = 0; r < 12; r++) {
    const uint8_t* s = B2B_SIGMA[r];
    G(0, 4, 8, 12, m[s[0]], m[s[1]]); G(1, 5, 9, 13, m[s[2]], m[s[3]]); G(2, 6, 10, 14, m[s[4]], m[s[5]]); G(3, 7, 11, 15, m[s[6]], m[s[7]]);
    G(0, 5, 10, 15, m[s[8]], m[s[9]]); G(1, 6, 11, 12, m[s[10]], m[s[11]]); G(2, 7, 8, 13, m[s[12]], m[s[13]]); G(3, 4, 9, 14, m[s[14]], m[s[15]]);
  }
#undef G
  for (int i = 0; i < 8; i++) c->h[i] ^= v[i] ^ v[i + 8];
}

void ok_blake2b_init(ok_blake2b_ctx* c, size_t outlen, const void* key, size_t keylen) {
  memset(c, 0, sizeof *c);
  for (int i = 0; i < 8; i++) c->h[i] = B2B_IV[i];
  c->h[0] ^= 0x01010000ULL ^ ((uint64_t)keylen << 8) ^ (uint64_t)outlen;
  c->outlen = outlen;
  if (keylen) {
    memcpy(c->buf, key, keylen); /* key block is zero padded to 128 bytes */
    c->fill = 128;
  }
}

void ok_blake2b_update(ok_blake2b_ctx* c, const void* data, size_t n) {
  const uint8_t* p = (const uint8_t*)data;
  while (n) {
    if (c->fill == 128) { /* buffer full and more input follows: not the last block */
      c->t[0] += 128;
      if (c->t[0] < 128) c->t[1]++;
      b2b_compress(c, c->buf, 0);
      c->fill = 0;
    }
    size_t take = 128 - c->fill;
    if (take > n) take = n;
    memcpy(c->buf + c->fill, p, take);
    c->fill += take; p += take; n -= take;
  }
}

void ok_blake2b_final(ok_blake2b_ctx* c, uint8_t* out) {
  c->t[0] += c->fill;
  if (c->t[0] < c->fill) c->t[1]++;
  memset(c->buf + c->fill, 0, 128 - c->fill);
  b2b_compress(c, c->buf, 1);
  uint8_t full[64];
  for (int i = 0; i < 8; i++)
    for (int j = 0; j < 8; j++) full[8 * i + j] = (uint8_t)(c->h[i] >> (8 * j));
  memcpy(out, full, c->outlen);
}

/* Reference: crypto/hashes/src/hashers.rs:84-90 — keyed BLAKE2b-256, key = ASCII domain. */
void ok_blake2b_keyed(const char* domain, const void* data, size_t n, uint8_t out[32]) {
  ok_blake2b_ctx c;
  ok_blake2b_init(&c, 32, domain, strlen(domain));
  ok_blake2b_update(&c, data, n);
  ok_blake2b_final(&c, out);
}

/* Unkeyed BLAKE2b-256: OpBlake2b (crypto/txscript/src/opcodes/mod.rs:738-744) and the
 * P2SH script hash (crypto/txscript/src/standard.rs:50-54). */
void ok_blake2b_256(const void* data, size_t n, uint8_t out[32]) {
  ok_blake2b_ctx c;
  ok_blake2b_init(&c, 32, NULL, 0);
  ok_blake2b_update(&c, data, n);
  ok_blake2b_final(&c, out);
}
