/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the shipped product path.
 *
 * "cpu_fast": a second, speed-oriented CPU port of the two verification routines, kept SEPARATE from the plain checker in
 * ok_secp256k1.c.  It exists so that the CPU baseline beside the GPU numbers is not a strawman: the plain checker is written for
 * obviousness (full reduction everywhere, no endomorphism, Jacobian tables) and runs ~3x slower per core than the C libsecp256k1 the
 * reference links (secp256k1-sys 0.10.1, Cargo.lock:5379-5396; not vendored under /root/reference, so it cannot be compiled here).
 * This file uses the same algorithmic ideas as that library - published ones, restated, no code copied:
 *   GLV endomorphism split of the variable-point scalar (128 doublings instead of 256), width-5 NAF over a per-call table of the odd
 *   multiples {1,3,...,15}P brought to a common Z on an isomorphic curve ("effective affine": mixed additions), lambda*P by one field
 *   multiplication per entry, the generator part through the 8-bit comb table of the checker, dedicated field squaring.
 * Verdicts (tri-state) must equal the plain checker's bit for bit: tests/test_oracle_secp.py compares the two on
 * random, corrupted and adversarial triples and on the BIP-340 vectors.  bench.py times THIS port (`cpu_baseline.kind = "port"`).
 */
#include "ok_oracle.h"
#include <string.h>

typedef unsigned __int128 u128;
typedef struct { uint64_t n[4]; } fe;
typedef struct { uint64_t n[4]; } sc;
typedef struct { fe x, y; int inf; } ge;
typedef struct { fe x, y, z; int inf; } gej;

#define FE_C 0x1000003D1ULL
static const uint64_t P_[4] = {0xFFFFFFFEFFFFFC2FULL, 0xFFFFFFFFFFFFFFFFULL, 0xFFFFFFFFFFFFFFFFULL, 0xFFFFFFFFFFFFFFFFULL};
static const uint64_t N_[4] = {0xBFD25E8CD0364141ULL, 0xBAAEDCE6AF48A03BULL, 0xFFFFFFFFFFFFFFFEULL, 0xFFFFFFFFFFFFFFFFULL};
static const uint64_t NC[3] = {0x402DA1732FC9BEBFULL, 0x4551231950B75FC4ULL, 1ULL};
static const uint64_t HALF_N[4] = {0xDFE92F46681B20A0ULL, 0x5D576E7357A4501DULL, 0xFFFFFFFFFFFFFFFFULL, 0x7FFFFFFFFFFFFFFFULL};
static const fe BETA = {{0xC1396C28719501EEULL, 0x9CF0497512F58995ULL, 0x6E64479EAC3434E9ULL, 0x7AE96A2B657C0710ULL}};
/* GLV lattice constants (rusty_kaspa_b200/csrc/kgv_secp.cuh carries the same numbers as 32-bit limbs; tools/derive_constants.py) */
static const uint64_t G1[4] = {0xE893209A45DBB031ULL, 0x3DAA8A1471E8CA7FULL, 0xE86C90E49284EB15ULL, 0x3086D221A7D46BCDULL};
static const uint64_t G2[4] = {0x1571B4AE8AC47F71ULL, 0x221208AC9DF506C6ULL, 0x6F547FA90ABFE4C4ULL, 0xE4437ED6010E8828ULL};
static const uint64_t A1[3] = {0xE86C90E49284EB15ULL, 0x3086D221A7D46BCDULL, 0};
static const uint64_t MB1[3] = {0x6F547FA90ABFE4C3ULL, 0xE4437ED6010E8828ULL, 0};
static const uint64_t A2[3] = {0x57C1108D9D44CFD8ULL, 0x14CA50F7A8E2F3F6ULL, 1};

/* ---------------------------------------------------------------- 256-bit helpers */
static inline int cmp4(const uint64_t* a, const uint64_t* b) { for (int i = 3; i >= 0; i--) if (a[i] != b[i]) return a[i] < b[i] ? -1 : 1; return 0; }
static inline uint64_t add4(uint64_t* r, const uint64_t* a, const uint64_t* b) { u128 c = 0; for (int i = 0; i < 4; i++) { c += (u128)a[i] + b[i]; r[i] = (uint64_t)c; c >>= 64; } return (uint64_t)c; }
static inline uint64_t sub4(uint64_t* r, const uint64_t* a, const uint64_t* b) { uint64_t br = 0; for (int i = 0; i < 4; i++) { u128 t = (u128)a[i] - b[i] - br; r[i] = (uint64_t)t; br = (uint64_t)(t >> 64) & 1; } return br; }
static inline int zero4(const uint64_t* a) { return (a[0] | a[1] | a[2] | a[3]) == 0; }
static void from_be(uint64_t r[4], const uint8_t b[32]) { for (int i = 0; i < 4; i++) { uint64_t w = 0; for (int k = 0; k < 8; k++) w = (w << 8) | b[8 * (3 - i) + k]; r[i] = w; } }

/* ---------------------------------------------------------------- field: fully reduced 4x64, dedicated squaring */
static inline void fe_reduce(fe* r, const uint64_t t[8]) {
  uint64_t a[4];
  u128 c = 0;
  for (int i = 0; i < 4; i++) { c += (u128)t[i] + (u128)t[4 + i] * FE_C; a[i] = (uint64_t)c; c >>= 64; }
  uint64_t top = (uint64_t)c;
  c = (u128)a[0] + (u128)top * FE_C; a[0] = (uint64_t)c; c >>= 64;
  for (int i = 1; i < 4; i++) { c += a[i]; a[i] = (uint64_t)c; c >>= 64; }
  if ((uint64_t)c) { c = (u128)a[0] + FE_C; a[0] = (uint64_t)c; c >>= 64; for (int i = 1; i < 4; i++) { c += a[i]; a[i] = (uint64_t)c; c >>= 64; } }
  if (cmp4(a, P_) >= 0) sub4(a, a, P_);
  memcpy(r->n, a, sizeof a);
}
static inline void fe_mul(fe* r, const fe* x, const fe* y) {
  const uint64_t *a = x->n, *b = y->n;
  uint64_t t[8];
  u128 c;
  /* operand scanning, fully unrolled by the compiler (-O3) */
  c = (u128)a[0] * b[0]; t[0] = (uint64_t)c; c >>= 64;
  c += (u128)a[0] * b[1]; t[1] = (uint64_t)c; c >>= 64;
  c += (u128)a[0] * b[2]; t[2] = (uint64_t)c; c >>= 64;
  c += (u128)a[0] * b[3]; t[3] = (uint64_t)c; t[4] = (uint64_t)(c >> 64);
  for (int i = 1; i < 4; i++) {
    c = 0;
    for (int j = 0; j < 4; j++) { c += (u128)a[i] * b[j] + t[i + j]; t[i + j] = (uint64_t)c; c >>= 64; }
    t[i + 4] = (uint64_t)c;
  }
  fe_reduce(r, t);
}
static inline void fe_sqr(fe* r, const fe* x) {
  const uint64_t* a = x->n;
  uint64_t t[8];
  u128 c;
  /* cross products once, doubled, plus the diagonal */
  c = (u128)a[0] * a[1]; t[1] = (uint64_t)c; c >>= 64;
  c += (u128)a[0] * a[2]; t[2] = (uint64_t)c; c >>= 64;
  c += (u128)a[0] * a[3]; t[3] = (uint64_t)c; t[4] = (uint64_t)(c >> 64);
  c = (u128)a[1] * a[2] + t[3]; t[3] = (uint64_t)c; c >>= 64;
  c += (u128)a[1] * a[3] + t[4]; t[4] = (uint64_t)c; t[5] = (uint64_t)(c >> 64);
  c = (u128)a[2] * a[3] + t[5]; t[5] = (uint64_t)c; t[6] = (uint64_t)(c >> 64);
  t[7] = t[6] >> 63;
  for (int i = 6; i > 1; i--) t[i] = (t[i] << 1) | (t[i - 1] >> 63);
  t[1] <<= 1;
  c = (u128)a[0] * a[0]; t[0] = (uint64_t)c; c >>= 64;
  c += t[1]; t[1] = (uint64_t)c; c >>= 64;
  for (int i = 1; i < 4; i++) {
    c += (u128)a[i] * a[i] + t[2 * i]; t[2 * i] = (uint64_t)c; c >>= 64;
    c += t[2 * i + 1]; t[2 * i + 1] = (uint64_t)c; c >>= 64;
  }
  fe_reduce(r, t);
}
static inline void fe_add(fe* r, const fe* a, const fe* b) { uint64_t t[4]; uint64_t c = add4(t, a->n, b->n); if (c || cmp4(t, P_) >= 0) sub4(t, t, P_); memcpy(r->n, t, sizeof t); }
static inline void fe_sub(fe* r, const fe* a, const fe* b) { uint64_t t[4]; if (sub4(t, a->n, b->n)) add4(t, t, P_); memcpy(r->n, t, sizeof t); }
static inline void fe_neg(fe* r, const fe* a) { if (zero4(a->n)) { *r = *a; return; } sub4(r->n, P_, a->n); }
static inline void fe_dbl(fe* r, const fe* a) { fe_add(r, a, a); }
static inline int fe_is_zero(const fe* a) { return zero4(a->n); }
static inline void fe_set_int(fe* r, uint64_t v) { r->n[0] = v; r->n[1] = r->n[2] = r->n[3] = 0; }
static void fe_sqr_n(fe* r, const fe* a, int n) { *r = *a; for (int i = 0; i < n; i++) fe_sqr(r, r); }
static void fe_pow_x223(fe* x223, fe* x22, fe* x2, fe* x3o, const fe* a) {
  fe x3, x6, x9, x11, x44, x88, t;
  fe_sqr(&t, a); fe_mul(x2, &t, a);
  fe_sqr(&t, x2); fe_mul(&x3, &t, a);
  fe_sqr_n(&t, &x3, 3); fe_mul(&x6, &t, &x3);
  fe_sqr_n(&t, &x6, 3); fe_mul(&x9, &t, &x3);
  fe_sqr_n(&t, &x9, 2); fe_mul(&x11, &t, x2);
  fe_sqr_n(&t, &x11, 11); fe_mul(x22, &t, &x11);
  fe_sqr_n(&t, x22, 22); fe_mul(&x44, &t, x22);
  fe_sqr_n(&t, &x44, 44); fe_mul(&x88, &t, &x44);
  fe_sqr_n(&t, &x88, 88); fe_mul(&t, &t, &x88);
  fe_sqr_n(&t, &t, 44); fe_mul(&t, &t, &x44);
  fe_sqr_n(&t, &t, 3); fe_mul(x223, &t, &x3);
  *x3o = x3;
}
static void fe_inv(fe* r, const fe* a) {
  fe x223, x22, x2, x3, t;
  fe_pow_x223(&x223, &x22, &x2, &x3, a);
  fe_sqr_n(&t, &x223, 23); fe_mul(&t, &t, &x22);
  fe_sqr_n(&t, &t, 5); fe_mul(&t, &t, a);
  fe_sqr_n(&t, &t, 3); fe_mul(&t, &t, &x2);
  fe_sqr_n(&t, &t, 2); fe_mul(r, &t, a);
}
static int fe_sqrt(fe* r, const fe* a) {
  fe x223, x22, x2, x3, t, chk;
  fe_pow_x223(&x223, &x22, &x2, &x3, a);
  fe_sqr_n(&t, &x223, 23); fe_mul(&t, &t, &x22);
  fe_sqr_n(&t, &t, 6); fe_mul(&t, &t, &x2);
  fe_sqr_n(&t, &t, 2);
  fe_sqr(&chk, &t);
  *r = t;
  return cmp4(chk.n, a->n) == 0;
}
static int ge_set_xo(ge* r, const fe* x, int odd) {
  fe x2, x3, seven, y;
  fe_sqr(&x2, x); fe_mul(&x3, &x2, x);
  fe_set_int(&seven, 7); fe_add(&x3, &x3, &seven);
  if (!fe_sqrt(&y, &x3)) return 0;
  if ((int)(y.n[0] & 1) != odd) fe_neg(&y, &y);
  r->x = *x; r->y = y; r->inf = 0;
  return 1;
}

/* ---------------------------------------------------------------- scalars */
static void sc_reduce512(sc* r, uint64_t* a) {
  /* fold the high 256 bits with 2^256 == NC (129 bits), three times */
  for (int round = 0; round < 3; round++) {
    uint64_t hi[4] = {a[4], a[5], a[6], a[7]};
    a[4] = a[5] = a[6] = a[7] = 0;
    for (int i = 0; i < 4; i++) {
      if (!hi[i]) continue;
      u128 c = 0;
      for (int j = 0; j < 3; j++) { c += (u128)hi[i] * NC[j] + a[i + j]; a[i + j] = (uint64_t)c; c >>= 64; }
      for (int k = i + 3; k < 8 && c; k++) { c += a[k]; a[k] = (uint64_t)c; c >>= 64; }
    }
  }
  if (cmp4(a, N_) >= 0) sub4(a, a, N_);
  memcpy(r->n, a, 32);
}
static void sc_mul(sc* r, const sc* x, const sc* y) {
  uint64_t t[8] = {0};
  for (int i = 0; i < 4; i++) { u128 c = 0; for (int j = 0; j < 4; j++) { c += (u128)x->n[i] * y->n[j] + t[i + j]; t[i + j] = (uint64_t)c; c >>= 64; } t[i + 4] = (uint64_t)c; }
  sc_reduce512(r, t);
}
static void sc_set_b32(sc* r, const uint8_t b[32], int* overflow) {
  from_be(r->n, b);
  int ov = cmp4(r->n, N_) >= 0;
  if (ov) sub4(r->n, r->n, N_);
  if (overflow) *overflow = ov;
}
static void sc_neg(sc* r, const sc* a) { if (zero4(a->n)) { *r = *a; return; } sub4(r->n, N_, a->n); }
static void sc_inv(sc* r, const sc* a) {
  static const uint64_t E[4] = {0xBFD25E8CD036413FULL, 0xBAAEDCE6AF48A03BULL, 0xFFFFFFFFFFFFFFFEULL, 0xFFFFFFFFFFFFFFFFULL};
  /* 4-bit fixed windows */
  sc tab[16];
  tab[1] = *a;
  sc_mul(&tab[2], a, a);
  for (int i = 3; i < 16; i++) sc_mul(&tab[i], &tab[i - 1], a);
  sc acc = tab[(E[3] >> 60) & 15];
  for (int nib = 62; nib >= 0; nib--) {
    for (int k = 0; k < 4; k++) sc_mul(&acc, &acc, &acc);
    unsigned d = (unsigned)((E[nib >> 4] >> (4 * (nib & 15))) & 15);
    if (d) sc_mul(&acc, &acc, &tab[d]);
  }
  *r = acc;
}

/* GLV: k = s1*|k1| + s2*|k2|*lambda (mod n), |k1|, |k2| < 2^129.  Integer arithmetic as in kgv_secp.cuh glv_split. */
static void mul_trunc3(uint64_t r[3], const uint64_t a[3], const uint64_t b[3]) {
  uint64_t t[3] = {0, 0, 0};
  for (int i = 0; i < 3; i++) { u128 c = 0; for (int j = 0; i + j < 3; j++) { c += (u128)a[i] * b[j] + t[i + j]; t[i + j] = (uint64_t)c; c >>= 64; } }
  memcpy(r, t, sizeof t);
}
static void sub3(uint64_t r[3], const uint64_t a[3], const uint64_t b[3]) { uint64_t br = 0; for (int i = 0; i < 3; i++) { u128 t = (u128)a[i] - b[i] - br; r[i] = (uint64_t)t; br = (uint64_t)(t >> 64) & 1; } }
static void neg3(uint64_t r[3]) { u128 c = 1; for (int i = 0; i < 3; i++) { c += (uint64_t)~r[i]; r[i] = (uint64_t)c; c >>= 64; } }
static void round_mul_shift384(uint64_t out[3], const uint64_t k[4], const uint64_t g[4]) {
  uint64_t t[8] = {0};
  for (int i = 0; i < 4; i++) { u128 c = 0; for (int j = 0; j < 4; j++) { c += (u128)k[i] * g[j] + t[i + j]; t[i + j] = (uint64_t)c; c >>= 64; } t[i + 4] = (uint64_t)c; }
  u128 c = t[5] >> 63;
  c += t[6]; out[0] = (uint64_t)c; c >>= 64;
  c += t[7]; out[1] = (uint64_t)c; out[2] = (uint64_t)(c >> 64);
}
static void glv_split(uint64_t k1[3], int* neg1, uint64_t k2[3], int* neg2, const sc* k) {
  uint64_t c1[3], c2[3], p1[3], p2[3], kk[3] = {k->n[0], k->n[1], k->n[2]};
  round_mul_shift384(c1, k->n, G1);
  round_mul_shift384(c2, k->n, G2);
  mul_trunc3(p1, c1, A1); mul_trunc3(p2, c2, A2);
  sub3(k1, kk, p1); sub3(k1, k1, p2);
  mul_trunc3(p1, c1, MB1); mul_trunc3(p2, c2, A1);
  sub3(k2, p1, p2);
  *neg1 = (int)(k1[2] >> 63); if (*neg1) neg3(k1);
  *neg2 = (int)(k2[2] >> 63); if (*neg2) neg3(k2);
}
/* width-5 NAF of a < 2^130 magnitude (3 limbs); returns the number of digits (<= 131) */
static int wnaf5(int8_t* out, const uint64_t m[3]) {
  uint64_t a[3] = {m[0], m[1], m[2]};
  int len = 0;
  while (a[0] | a[1] | a[2]) {
    int d = 0;
    if (a[0] & 1) {
      d = (int)(a[0] & 31);
      if (d >= 16) d -= 32;
      if (d > 0) { uint64_t b = (uint64_t)d; for (int i = 0; i < 3 && b; i++) { uint64_t o = a[i]; a[i] = o - b; b = o < b; } }
      else { uint64_t c = (uint64_t)(-d); for (int i = 0; i < 3 && c; i++) { uint64_t o = a[i]; a[i] = o + c; c = a[i] < o; } }
    }
    out[len++] = (int8_t)d;
    a[0] = (a[0] >> 1) | (a[1] << 63); a[1] = (a[1] >> 1) | (a[2] << 63); a[2] >>= 1;
  }
  return len;
}

/* ---------------------------------------------------------------- group law (a = 0) */
static void gej_double(gej* r, const gej* a) {
  if (a->inf || fe_is_zero(&a->y)) { memset(r, 0, sizeof *r); r->inf = 1; return; }
  fe A, B, C, D, E, F, t, z3, x3, y3;
  fe_sqr(&A, &a->x); fe_sqr(&B, &a->y); fe_sqr(&C, &B);
  fe_add(&t, &a->x, &B); fe_sqr(&t, &t); fe_sub(&t, &t, &A); fe_sub(&t, &t, &C); fe_dbl(&D, &t);
  fe_dbl(&E, &A); fe_add(&E, &E, &A);
  fe_sqr(&F, &E);
  fe_mul(&z3, &a->y, &a->z); fe_dbl(&z3, &z3);
  fe_sub(&x3, &F, &D); fe_sub(&x3, &x3, &D);
  fe_sub(&t, &D, &x3); fe_mul(&y3, &E, &t);
  fe_dbl(&C, &C); fe_dbl(&C, &C); fe_dbl(&C, &C);
  fe_sub(&y3, &y3, &C);
  r->x = x3; r->y = y3; r->z = z3; r->inf = 0;
}
/* r = a + (bx, by) affine, never infinity; *hout (may be NULL) receives the factor Z was multiplied by */
static void gej_add_ge(gej* r, const gej* a, const fe* bx, const fe* by, fe* hout) {
  if (a->inf) { r->x = *bx; r->y = *by; fe_set_int(&r->z, 1); r->inf = 0; if (hout) fe_set_int(hout, 1); return; }
  fe z1z1, u2, s2, h, rr, t;
  fe_sqr(&z1z1, &a->z);
  fe_mul(&u2, bx, &z1z1);
  fe_mul(&t, &a->z, &z1z1); fe_mul(&s2, by, &t);
  fe_sub(&h, &u2, &a->x); fe_sub(&rr, &s2, &a->y);
  if (fe_is_zero(&h)) {
    if (hout) fe_set_int(hout, 1);
    if (fe_is_zero(&rr)) gej_double(r, a); else { memset(r, 0, sizeof *r); r->inf = 1; }
    return;
  }
  if (hout) *hout = h;
  fe hh, hhh, v, x3, y3, z3;
  fe_sqr(&hh, &h); fe_mul(&hhh, &hh, &h); fe_mul(&v, &a->x, &hh);
  fe_sqr(&x3, &rr); fe_sub(&x3, &x3, &hhh); fe_sub(&x3, &x3, &v); fe_sub(&x3, &x3, &v);
  fe_sub(&t, &v, &x3); fe_mul(&y3, &rr, &t); fe_mul(&t, &a->y, &hhh); fe_sub(&y3, &y3, &t);
  fe_mul(&z3, &a->z, &h);
  r->x = x3; r->y = y3; r->z = z3; r->inf = 0;
}
static void gej_add(gej* r, const gej* a, const gej* b) {
  if (a->inf) { *r = *b; return; }
  if (b->inf) { *r = *a; return; }
  fe z1z1, z2z2, u1, u2, s1, s2, h, rr, t;
  fe_sqr(&z1z1, &a->z); fe_sqr(&z2z2, &b->z);
  fe_mul(&u1, &a->x, &z2z2); fe_mul(&u2, &b->x, &z1z1);
  fe_mul(&t, &b->z, &z2z2); fe_mul(&s1, &a->y, &t);
  fe_mul(&t, &a->z, &z1z1); fe_mul(&s2, &b->y, &t);
  fe_sub(&h, &u2, &u1); fe_sub(&rr, &s2, &s1);
  if (fe_is_zero(&h)) { if (fe_is_zero(&rr)) gej_double(r, a); else { memset(r, 0, sizeof *r); r->inf = 1; } return; }
  fe hh, hhh, v, x3, y3, z3;
  fe_sqr(&hh, &h); fe_mul(&hhh, &hh, &h); fe_mul(&v, &u1, &hh);
  fe_sqr(&x3, &rr); fe_sub(&x3, &x3, &hhh); fe_sub(&x3, &x3, &v); fe_sub(&x3, &x3, &v);
  fe_sub(&t, &v, &x3); fe_mul(&y3, &rr, &t); fe_mul(&t, &s1, &hhh); fe_sub(&y3, &y3, &t);
  fe_mul(&z3, &a->z, &b->z); fe_mul(&z3, &z3, &h);
  r->x = x3; r->y = y3; r->z = z3; r->inf = 0;
}

/* ---------------------------------------------------------------- generator comb: 8-bit windows, built once (Montgomery batch inversion) */
static ge (*GT)[255];
static int gt_ready;
#include <pthread.h>
#include <stdlib.h>
static pthread_once_t gt_once = PTHREAD_ONCE_INIT;
static void build_gt(void) {
  static const fe GX = {{0x59F2815B16F81798ULL, 0x029BFCDB2DCE28D9ULL, 0x55A06295CE870B07ULL, 0x79BE667EF9DCBBACULL}};
  static const fe GY = {{0x9C47D08FFB10D4B8ULL, 0xFD17B448A6855419ULL, 0x5DA4FBFC0E1108A8ULL, 0x483ADA7726A3C465ULL}};
  GT = malloc(sizeof(ge[255]) * 32);
  size_t N = 32 * 255;
  gej* tmp = malloc(sizeof(gej) * N);
  fe bx = GX, by = GY;
  for (int i = 0; i < 32; i++) {
    gej acc; memset(&acc, 0, sizeof acc); acc.inf = 1;
    for (int j = 1; j <= 255; j++) { gej_add_ge(&acc, &acc, &bx, &by, NULL); tmp[i * 255 + j - 1] = acc; }
    gej nb; gej_add_ge(&nb, &acc, &bx, &by, NULL); /* 256 * base */
    fe zi, zi2, zi3; fe_inv(&zi, &nb.z); fe_sqr(&zi2, &zi); fe_mul(&zi3, &zi2, &zi);
    fe_mul(&bx, &nb.x, &zi2); fe_mul(&by, &nb.y, &zi3);
  }
  fe* pre = malloc(sizeof(fe) * N);
  fe run; fe_set_int(&run, 1);
  for (size_t k = 0; k < N; k++) { pre[k] = run; fe_mul(&run, &run, &tmp[k].z); }
  fe inv; fe_inv(&inv, &run);
  for (size_t k = N; k-- > 0;) {
    fe zi, zi2, zi3;
    fe_mul(&zi, &inv, &pre[k]); fe_mul(&inv, &inv, &tmp[k].z);
    fe_sqr(&zi2, &zi); fe_mul(&zi3, &zi2, &zi);
    ge* o = &GT[k / 255][k % 255];
    fe_mul(&o->x, &tmp[k].x, &zi2); fe_mul(&o->y, &tmp[k].y, &zi3); o->inf = 0;
  }
  free(pre); free(tmp);
  gt_ready = 1;
}

/* R = na*P + ng*G.  Returns through (R, zs): R lives on secp256k1 directly (the P part is mapped back before the G part is added). */
static void ecmult_fast(gej* out, const ge* P, const sc* na, const sc* ng) {
  pthread_once(&gt_once, build_gt);
  gej acc; memset(&acc, 0, sizeof acc); acc.inf = 1;
  if (!zero4(na->n)) {
    /* odd multiples 1,3,..,15 of P on the curve where D = 2P is affine, then brought to the Z of the last one */
    gej d, pj; pj.x = P->x; pj.y = P->y; fe_set_int(&pj.z, 1); pj.inf = 0;
    gej_double(&d, &pj);
    fe zd2, zd3, tx[8], ty[8], H[7];
    fe_sqr(&zd2, &d.z); fe_mul(&zd3, &zd2, &d.z);
    gej t; fe_mul(&t.x, &P->x, &zd2); fe_mul(&t.y, &P->y, &zd3); fe_set_int(&t.z, 1); t.inf = 0;
    tx[0] = t.x; ty[0] = t.y;
    for (int j = 1; j < 8; j++) { gej_add_ge(&t, &t, &d.x, &d.y, &H[j - 1]); tx[j] = t.x; ty[j] = t.y; }
    fe accz = H[6];
    for (int j = 6; j >= 0; j--) {
      if (j < 6) fe_mul(&accz, &accz, &H[j]);
      fe a2, a3; fe_sqr(&a2, &accz); fe_mul(&a3, &a2, &accz);
      fe_mul(&tx[j], &tx[j], &a2); fe_mul(&ty[j], &ty[j], &a3);
    }
    fe zs; fe_mul(&zs, &t.z, &d.z);  /* true Z = Z' * zs */
    fe lx[8];
    for (int j = 0; j < 8; j++) fe_mul(&lx[j], &tx[j], &BETA);
    uint64_t m1[3], m2[3];
    int n1, n2;
    glv_split(m1, &n1, m2, &n2, na);
    int8_t w1[136], w2[136];
    memset(w1, 0, sizeof w1); memset(w2, 0, sizeof w2);
    int l1 = wnaf5(w1, m1), l2 = wnaf5(w2, m2);
    int len = l1 > l2 ? l1 : l2;
    for (int i = len - 1; i >= 0; i--) {
      gej_double(&acc, &acc);
      int dgt = w1[i];
      if (dgt) {
        int neg = (dgt < 0) != (n1 != 0);
        int idx = ((dgt < 0 ? -dgt : dgt) - 1) >> 1;
        fe y = ty[idx]; if (neg) fe_neg(&y, &y);
        gej_add_ge(&acc, &acc, &tx[idx], &y, NULL);
      }
      dgt = w2[i];
      if (dgt) {
        int neg = (dgt < 0) != (n2 != 0);
        int idx = ((dgt < 0 ? -dgt : dgt) - 1) >> 1;
        fe y = ty[idx]; if (neg) fe_neg(&y, &y);
        gej_add_ge(&acc, &acc, &lx[idx], &y, NULL);
      }
    }
    if (!acc.inf) fe_mul(&acc.z, &acc.z, &zs);  /* back onto secp256k1 */
  }
  gej g; memset(&g, 0, sizeof g); g.inf = 1;
  for (int i = 0; i < 32; i++) {
    unsigned dg = (unsigned)((ng->n[i >> 3] >> (8 * (i & 7))) & 0xff);
    if (dg) gej_add_ge(&g, &g, &GT[i][dg - 1].x, &GT[i][dg - 1].y, NULL);
  }
  gej_add(out, &acc, &g);
}

/* ---------------------------------------------------------------- verification (same tri-state rules as ok_secp256k1.c) */
int ok_schnorr_verify_fast(const uint8_t pk32[32], const uint8_t msg32[32], const uint8_t sig64[64]) {
  fe px, rx;
  ge P;
  from_be(px.n, pk32);
  if (cmp4(px.n, P_) >= 0) return OK_SIG_PK_PARSE_ERR;
  if (!ge_set_xo(&P, &px, 0)) return OK_SIG_PK_PARSE_ERR;
  from_be(rx.n, sig64);
  if (cmp4(rx.n, P_) >= 0) return OK_SIG_INVALID;
  sc s, e;
  int ov;
  sc_set_b32(&s, sig64 + 32, &ov);
  if (ov) return OK_SIG_INVALID;
  /* BIP-340 challenge: tagged hash with the tag midstate recomputed per call (cheap next to the curve arithmetic) */
  uint8_t th[32], eh[32];
  ok_sha256("BIP0340/challenge", 17, th);
  ok_sha256_ctx h;
  ok_sha256_init(&h);
  ok_sha256_update(&h, th, 32); ok_sha256_update(&h, th, 32);
  ok_sha256_update(&h, sig64, 32); ok_sha256_update(&h, pk32, 32); ok_sha256_update(&h, msg32, 32);
  ok_sha256_final(&h, eh);
  sc_set_b32(&e, eh, NULL);
  sc_neg(&e, &e);
  gej R;
  ecmult_fast(&R, &P, &e, &s);
  if (R.inf) return OK_SIG_INVALID;
  fe zi, zi2, zi3, ax, ay;
  fe_inv(&zi, &R.z); fe_sqr(&zi2, &zi); fe_mul(&zi3, &zi2, &zi);
  fe_mul(&ax, &R.x, &zi2); fe_mul(&ay, &R.y, &zi3);
  if (ay.n[0] & 1) return OK_SIG_INVALID;
  return cmp4(ax.n, rx.n) == 0 ? OK_SIG_VALID : OK_SIG_INVALID;
}

int ok_ecdsa_verify_fast(const uint8_t pk33[33], const uint8_t msg32[32], const uint8_t sig64[64]) {
  fe qx;
  ge Q;
  if (pk33[0] != 0x02 && pk33[0] != 0x03) return OK_SIG_PK_PARSE_ERR;
  from_be(qx.n, pk33 + 1);
  if (cmp4(qx.n, P_) >= 0) return OK_SIG_PK_PARSE_ERR;
  if (!ge_set_xo(&Q, &qx, pk33[0] == 0x03)) return OK_SIG_PK_PARSE_ERR;
  sc r, s, m;
  int ovr, ovs;
  sc_set_b32(&r, sig64, &ovr);
  sc_set_b32(&s, sig64 + 32, &ovs);
  if (ovr || ovs) return OK_SIG_SIG_PARSE_ERR;
  sc_set_b32(&m, msg32, NULL);
  if (cmp4(s.n, HALF_N) > 0) return OK_SIG_INVALID;
  if (zero4(r.n) || zero4(s.n)) return OK_SIG_INVALID;
  sc sn, u1, u2;
  sc_inv(&sn, &s);
  sc_mul(&u1, &sn, &m);
  sc_mul(&u2, &sn, &r);
  gej R;
  ecmult_fast(&R, &Q, &u2, &u1);
  if (R.inf) return OK_SIG_INVALID;
  /* x(R) mod n == r, checked projectively: X == r*Z^2, or (r + n < p and X == (r+n)*Z^2) */
  fe z2, t, rf;
  fe_sqr(&z2, &R.z);
  memcpy(rf.n, r.n, 32);
  fe_mul(&t, &rf, &z2);
  if (cmp4(t.n, R.x.n) == 0) return OK_SIG_VALID;
  uint64_t rn[4];
  if (add4(rn, r.n, N_) || cmp4(rn, P_) >= 0) return OK_SIG_INVALID;
  memcpy(rf.n, rn, 32);
  fe_mul(&t, &rf, &z2);
  return cmp4(t.n, R.x.n) == 0 ? OK_SIG_VALID : OK_SIG_INVALID;
}

typedef struct { const uint8_t *pk, *msg, *sig; uint8_t* st; size_t lo, hi; int ecdsa; } fjob;
static void* fworker(void* a) {
  fjob* j = (fjob*)a;
  for (size_t i = j->lo; i < j->hi; i++)
    j->st[i] = (uint8_t)(j->ecdsa ? ok_ecdsa_verify_fast(j->pk + 33 * i, j->msg + 32 * i, j->sig + 64 * i) : ok_schnorr_verify_fast(j->pk + 32 * i, j->msg + 32 * i, j->sig + 64 * i));
  return NULL;
}
static void run_fast(const uint8_t* pk, const uint8_t* msg, const uint8_t* sig, size_t n, uint8_t* st, int nthreads, int ecdsa) {
  pthread_once(&gt_once, build_gt);
  if (nthreads < 1) nthreads = 1;
  if ((size_t)nthreads > n) nthreads = n ? (int)n : 1;
  pthread_t* th = malloc(sizeof(pthread_t) * (size_t)nthreads);
  fjob* jobs = malloc(sizeof(fjob) * (size_t)nthreads);
  for (int t = 0; t < nthreads; t++) {
    jobs[t] = (fjob){pk, msg, sig, st, n * (size_t)t / (size_t)nthreads, n * ((size_t)t + 1) / (size_t)nthreads, ecdsa};
    if (t + 1 < nthreads) pthread_create(&th[t], NULL, fworker, &jobs[t]);
  }
  fworker(&jobs[nthreads - 1]);
  for (int t = 0; t + 1 < nthreads; t++) pthread_join(th[t], NULL);
  free(th); free(jobs);
}
void ok_schnorr_verify_batch_fast(const uint8_t* pk32, const uint8_t* msg32, const uint8_t* sig64, size_t n, uint8_t* status, int nthreads) { run_fast(pk32, msg32, sig64, n, status, nthreads, 0); }
void ok_ecdsa_verify_batch_fast(const uint8_t* pk33, const uint8_t* msg32, const uint8_t* sig64, size_t n, uint8_t* status, int nthreads) { run_fast(pk33, msg32, sig64, n, status, nthreads, 1); }
