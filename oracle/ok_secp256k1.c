/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the shipped product path.
 *
 * secp256k1 field / scalar / group arithmetic, BIP-340 Schnorr verification and ECDSA
 * verification with the exact accept/reject/parse-error semantics of the library the
 * reference calls: `secp256k1 0.29.1` -> `secp256k1-sys 0.10.1` (bundled libsecp256k1;
 * Cargo.lock:5379-5396).  That source is NOT vendored under /root/reference; this file
 * restates the published algorithms (SEC 2 curve parameters, BIP-340, SEC 1 ECDSA) and
 * follows the reference's own call sites:
 *   crypto/txscript/src/lib.rs:582  XOnlyPublicKey::from_slice   -> x < p and x^3+7 is a square
 *   crypto/txscript/src/lib.rs:583  schnorr::Signature::from_slice -> length check only
 *   crypto/txscript/src/lib.rs:593  sig.verify(&msg,&pk)          -> BIP-340 verify
 *   crypto/txscript/src/lib.rs:618  PublicKey::from_slice (33 B)  -> tag 02/03, x < p, on curve
 *   crypto/txscript/src/lib.rs:619  ecdsa::Signature::from_compact -> r < n and s < n
 *   crypto/txscript/src/lib.rs:628  sig.verify(&msg,&pk)          -> ECDSA verify, low-S required
 * Representation: 4 x 64-bit little-endian limbs, always fully reduced; products via
 * unsigned __int128.  (The CUDA kernels use an unrelated 8 x 32-bit layout.)
 */
#include "ok_oracle.h"
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned __int128 u128;
typedef struct { uint64_t n[4]; } fe; /* field element mod p, < p */
typedef struct { uint64_t n[4]; } sc; /* scalar mod n, < n */
typedef struct { fe x, y; int inf; } ge;
typedef struct { fe x, y, z; int inf; } gej;

/* p = 2^256 - 2^32 - 977 */
static const fe FE_P = {{0xFFFFFFFEFFFFFC2FULL, 0xFFFFFFFFFFFFFFFFULL, 0xFFFFFFFFFFFFFFFFULL, 0xFFFFFFFFFFFFFFFFULL}};
#define FE_C 0x1000003D1ULL
/* n = group order */
static const sc SC_N = {{0xBFD25E8CD0364141ULL, 0xBAAEDCE6AF48A03BULL, 0xFFFFFFFFFFFFFFFEULL, 0xFFFFFFFFFFFFFFFFULL}};
static const uint64_t SC_NC[3] = {0x402DA1732FC9BEBFULL, 0x4551231950B75FC4ULL, 1ULL}; /* 2^256 - n */
static const sc SC_HALF_N = {{0xDFE92F46681B20A0ULL, 0x5D576E7357A4501DULL, 0xFFFFFFFFFFFFFFFFULL, 0x7FFFFFFFFFFFFFFFULL}}; /* (n-1)/2 */
static const ge GE_G = {{{0x59F2815B16F81798ULL, 0x029BFCDB2DCE28D9ULL, 0x55A06295CE870B07ULL, 0x79BE667EF9DCBBACULL}},
                        {{0x9C47D08FFB10D4B8ULL, 0xFD17B448A6855419ULL, 0x5DA4FBFC0E1108A8ULL, 0x483ADA7726A3C465ULL}},
                        0};

/* ------------------------------------------------------------ 256-bit helpers */
static int u256_cmp(const uint64_t a[4], const uint64_t b[4]) {
  for (int i = 3; i >= 0; i--) {
    if (a[i] < b[i]) return -1;
    if (a[i] > b[i]) return 1;
  }
  return 0;
}
static int u256_is_zero(const uint64_t a[4]) { return (a[0] | a[1] | a[2] | a[3]) == 0; }
static uint64_t u256_add(uint64_t r[4], const uint64_t a[4], const uint64_t b[4]) {
  u128 c = 0;
  for (int i = 0; i < 4; i++) { c += (u128)a[i] + b[i]; r[i] = (uint64_t)c; c >>= 64; }
  return (uint64_t)c;
}
static uint64_t u256_sub(uint64_t r[4], const uint64_t a[4], const uint64_t b[4]) {
  uint64_t borrow = 0;
  for (int i = 0; i < 4; i++) {
    uint64_t d = a[i] - b[i];
    uint64_t b1 = a[i] < b[i];
    uint64_t d2 = d - borrow;
    uint64_t b2 = d < borrow;
    r[i] = d2;
    borrow = b1 | b2;
  }
  return borrow;
}
static void u256_from_be(uint64_t r[4], const uint8_t b[32]) {
  for (int i = 0; i < 4; i++) {
    uint64_t w = 0;
    for (int j = 0; j < 8; j++) w = (w << 8) | b[8 * (3 - i) + j];
    r[i] = w;
  }
}
static void u256_to_be(uint8_t b[32], const uint64_t a[4]) {
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 8; j++) b[8 * (3 - i) + j] = (uint8_t)(a[i] >> (56 - 8 * j));
}
static void u256_mul_wide(uint64_t t[8], const uint64_t a[4], const uint64_t b[4]) {
  memset(t, 0, 8 * sizeof(uint64_t));
  for (int i = 0; i < 4; i++) {
    u128 c = 0;
    for (int j = 0; j < 4; j++) {
      c += (u128)a[i] * b[j] + t[i + j];
      t[i + j] = (uint64_t)c;
      c >>= 64;
    }
    t[i + 4] = (uint64_t)c;
  }
}

/* ------------------------------------------------------------ field */
static void fe_reduce_wide(fe* r, const uint64_t t[8]) {
  /* 2^256 == C (mod p): fold the high half twice, then one conditional subtract. */
  uint64_t a[4];
  u128 c = 0;
  for (int i = 0; i < 4; i++) { c += (u128)t[i] + (u128)t[4 + i] * FE_C; a[i] = (uint64_t)c; c >>= 64; }
  uint64_t top = (uint64_t)c; /* < 2^34 */
  c = (u128)a[0] + (u128)top * FE_C; a[0] = (uint64_t)c; c >>= 64;
  for (int i = 1; i < 4; i++) { c += a[i]; a[i] = (uint64_t)c; c >>= 64; }
  if ((uint64_t)c) { /* wrapped past 2^256: remaining value is tiny, add C once more */
    c = (u128)a[0] + FE_C; a[0] = (uint64_t)c; c >>= 64;
    for (int i = 1; i < 4; i++) { c += a[i]; a[i] = (uint64_t)c; c >>= 64; }
  }
  if (u256_cmp(a, FE_P.n) >= 0) u256_sub(a, a, FE_P.n);
  memcpy(r->n, a, sizeof a);
}
static void fe_mul(fe* r, const fe* a, const fe* b) { uint64_t t[8]; u256_mul_wide(t, a->n, b->n); fe_reduce_wide(r, t); }
static void fe_sqr(fe* r, const fe* a) { fe_mul(r, a, a); }
static void fe_add(fe* r, const fe* a, const fe* b) {
  uint64_t t[4];
  uint64_t carry = u256_add(t, a->n, b->n);
  if (carry || u256_cmp(t, FE_P.n) >= 0) u256_sub(t, t, FE_P.n); /* carry: t - p wraps to the right value */
  memcpy(r->n, t, sizeof t);
}
static void fe_sub(fe* r, const fe* a, const fe* b) {
  uint64_t t[4];
  if (u256_sub(t, a->n, b->n)) u256_add(t, t, FE_P.n);
  memcpy(r->n, t, sizeof t);
}
static void fe_neg(fe* r, const fe* a) {
  if (u256_is_zero(a->n)) { *r = *a; return; }
  u256_sub(r->n, FE_P.n, a->n);
}
static void fe_mul_small(fe* r, const fe* a, unsigned k) { /* k in 2..8 */
  fe acc = *a;
  for (unsigned i = 1; i < k; i++) fe_add(&acc, &acc, a);
  *r = acc;
}
static int fe_is_zero(const fe* a) { return u256_is_zero(a->n); }
static int fe_equal(const fe* a, const fe* b) { return u256_cmp(a->n, b->n) == 0; }
static int fe_is_odd(const fe* a) { return (int)(a->n[0] & 1); }
static void fe_set_int(fe* r, uint64_t v) { r->n[0] = v; r->n[1] = r->n[2] = r->n[3] = 0; }
/* returns 0 if the 32-byte big-endian value is >= p */
static int fe_set_b32_limit(fe* r, const uint8_t b[32]) { u256_from_be(r->n, b); return u256_cmp(r->n, FE_P.n) < 0; }
static void fe_get_b32(uint8_t b[32], const fe* a) { u256_to_be(b, a->n); }
static void fe_sqr_n(fe* r, const fe* a, int n) { *r = *a; for (int i = 0; i < n; i++) fe_sqr(r, r); }

/* a^(2^223-1) and friends: shared prefix of the addition chains for p-2 and (p+1)/4.
 * p = 2^256 - 2^32 - 977; the exponents are (from the top) 223 ones, a zero, 22 ones, then a short tail. */
static void fe_pow_x223(fe* x223, fe* x22, fe* x2, fe* x3out, const fe* a) {
  fe x3, x6, x9, x11, x44, x88, x176, x220, t;
  fe_sqr(&t, a); fe_mul(x2, &t, a);
  fe_sqr(&t, x2); fe_mul(&x3, &t, a);
  fe_sqr_n(&t, &x3, 3); fe_mul(&x6, &t, &x3);
  fe_sqr_n(&t, &x6, 3); fe_mul(&x9, &t, &x3);
  fe_sqr_n(&t, &x9, 2); fe_mul(&x11, &t, x2);
  fe_sqr_n(&t, &x11, 11); fe_mul(x22, &t, &x11);
  fe_sqr_n(&t, x22, 22); fe_mul(&x44, &t, x22);
  fe_sqr_n(&t, &x44, 44); fe_mul(&x88, &t, &x44);
  fe_sqr_n(&t, &x88, 88); fe_mul(&x176, &t, &x88);
  fe_sqr_n(&t, &x176, 44); fe_mul(&x220, &t, &x44);
  fe_sqr_n(&t, &x220, 3); fe_mul(x223, &t, &x3);
  *x3out = x3;
}
static void fe_inv(fe* r, const fe* a) { /* a^(p-2); 0 -> 0 */
  fe x223, x22, x2, x3, t;
  fe_pow_x223(&x223, &x22, &x2, &x3, a);
  fe_sqr_n(&t, &x223, 23); fe_mul(&t, &t, &x22);
  fe_sqr_n(&t, &t, 5); fe_mul(&t, &t, a);
  fe_sqr_n(&t, &t, 3); fe_mul(&t, &t, &x2);
  fe_sqr_n(&t, &t, 2); fe_mul(r, &t, a);
}
/* r = a^((p+1)/4); returns 1 iff r^2 == a (a is a quadratic residue) */
static int fe_sqrt(fe* r, const fe* a) {
  fe x223, x22, x2, x3, t, chk;
  fe_pow_x223(&x223, &x22, &x2, &x3, a);
  fe_sqr_n(&t, &x223, 23); fe_mul(&t, &t, &x22);
  fe_sqr_n(&t, &t, 6); fe_mul(&t, &t, &x2);
  fe_sqr_n(&t, &t, 2);
  fe_sqr(&chk, &t);
  *r = t;
  return fe_equal(&chk, a);
}

/* ------------------------------------------------------------ scalar */
static void sc_reduce_limbs(sc* r, uint64_t* a /* 8 limbs, destroyed */) {
  /* 2^256 == NC (mod n), NC is 129 bits: fold the part above 2^256 until it is gone. */
  for (;;) {
    if ((a[4] | a[5] | a[6] | a[7]) == 0) break;
    uint64_t hi[4] = {a[4], a[5], a[6], a[7]};
    uint64_t acc[8] = {a[0], a[1], a[2], a[3], 0, 0, 0, 0};
    for (int i = 0; i < 4; i++) {
      if (!hi[i]) continue;
      u128 c = 0;
      int k = i;
      for (int j = 0; j < 3; j++, k++) { c += (u128)hi[i] * SC_NC[j] + acc[k]; acc[k] = (uint64_t)c; c >>= 64; }
      for (; c && k < 8; k++) { c += acc[k]; acc[k] = (uint64_t)c; c >>= 64; }
    }
    memcpy(a, acc, sizeof acc);
  }
  if (u256_cmp(a, SC_N.n) >= 0) u256_sub(a, a, SC_N.n);
  memcpy(r->n, a, 4 * sizeof(uint64_t));
}
static void sc_mul(sc* r, const sc* a, const sc* b) { uint64_t t[8]; u256_mul_wide(t, a->n, b->n); sc_reduce_limbs(r, t); }
static void sc_add(sc* r, const sc* a, const sc* b) {
  uint64_t t[8] = {0};
  t[4] = u256_add(t, a->n, b->n);
  sc_reduce_limbs(r, t);
}
static void sc_neg(sc* r, const sc* a) {
  if (u256_is_zero(a->n)) { *r = *a; return; }
  u256_sub(r->n, SC_N.n, a->n);
}
static int sc_is_zero(const sc* a) { return u256_is_zero(a->n); }
static int sc_is_high(const sc* a) { return u256_cmp(a->n, SC_HALF_N.n) > 0; }
/* sets r = b mod n; *overflow = (b >= n) */
static void sc_set_b32(sc* r, const uint8_t b[32], int* overflow) {
  uint64_t t[8] = {0};
  u256_from_be(t, b);
  int ov = u256_cmp(t, SC_N.n) >= 0;
  if (overflow) *overflow = ov;
  sc_reduce_limbs(r, t);
}
static void sc_get_b32(uint8_t b[32], const sc* a) { u256_to_be(b, a->n); }
static void sc_inv(sc* r, const sc* a) { /* a^(n-2), square and multiply */
  uint64_t e[4];
  uint64_t two[4] = {2, 0, 0, 0};
  u256_sub(e, SC_N.n, two);
  sc acc = {{1, 0, 0, 0}};
  for (int i = 255; i >= 0; i--) {
    sc_mul(&acc, &acc, &acc);
    if ((e[i >> 6] >> (i & 63)) & 1) sc_mul(&acc, &acc, a);
  }
  *r = acc;
}

/* ------------------------------------------------------------ group */
static void gej_set_inf(gej* r) { memset(r, 0, sizeof *r); r->inf = 1; }
static void gej_set_ge(gej* r, const ge* a) { r->x = a->x; r->y = a->y; fe_set_int(&r->z, 1); r->inf = a->inf; }

static void gej_double(gej* r, const gej* a) {
  if (a->inf || fe_is_zero(&a->y)) { gej_set_inf(r); return; }
  fe A, B, C, D, E, F, t;
  fe_sqr(&A, &a->x);
  fe_sqr(&B, &a->y);
  fe_sqr(&C, &B);
  fe_add(&t, &a->x, &B); fe_sqr(&t, &t); fe_sub(&t, &t, &A); fe_sub(&t, &t, &C); fe_add(&D, &t, &t);
  fe_mul_small(&E, &A, 3);
  fe_sqr(&F, &E);
  fe z3; fe_mul(&z3, &a->y, &a->z); fe_add(&z3, &z3, &z3);
  fe x3; fe_sub(&x3, &F, &D); fe_sub(&x3, &x3, &D);
  fe y3; fe_sub(&t, &D, &x3); fe_mul(&y3, &E, &t); fe_mul_small(&t, &C, 8); fe_sub(&y3, &y3, &t);
  r->x = x3; r->y = y3; r->z = z3; r->inf = 0;
}

/* r = a + b with b given in Jacobian coordinates; handles every special case */
static void gej_add(gej* r, const gej* a, const gej* b) {
  if (a->inf) { *r = *b; return; }
  if (b->inf) { *r = *a; return; }
  fe z1z1, z2z2, u1, u2, s1, s2, h, rr, t;
  fe_sqr(&z1z1, &a->z); fe_sqr(&z2z2, &b->z);
  fe_mul(&u1, &a->x, &z2z2); fe_mul(&u2, &b->x, &z1z1);
  fe_mul(&t, &b->z, &z2z2); fe_mul(&s1, &a->y, &t);
  fe_mul(&t, &a->z, &z1z1); fe_mul(&s2, &b->y, &t);
  fe_sub(&h, &u2, &u1); fe_sub(&rr, &s2, &s1);
  if (fe_is_zero(&h)) {
    if (fe_is_zero(&rr)) gej_double(r, a); else gej_set_inf(r);
    return;
  }
  fe hh, hhh, v, x3, y3, z3;
  fe_sqr(&hh, &h); fe_mul(&hhh, &hh, &h); fe_mul(&v, &u1, &hh);
  fe_sqr(&x3, &rr); fe_sub(&x3, &x3, &hhh); fe_sub(&x3, &x3, &v); fe_sub(&x3, &x3, &v);
  fe_sub(&t, &v, &x3); fe_mul(&y3, &rr, &t); fe_mul(&t, &s1, &hhh); fe_sub(&y3, &y3, &t);
  fe_mul(&z3, &a->z, &b->z); fe_mul(&z3, &z3, &h);
  r->x = x3; r->y = y3; r->z = z3; r->inf = 0;
}

/* r = a + b with b affine */
static void gej_add_ge(gej* r, const gej* a, const ge* b) {
  if (b->inf) { *r = *a; return; }
  if (a->inf) { gej_set_ge(r, b); return; }
  fe z1z1, u2, s2, h, rr, t;
  fe_sqr(&z1z1, &a->z);
  fe_mul(&u2, &b->x, &z1z1);
  fe_mul(&t, &a->z, &z1z1); fe_mul(&s2, &b->y, &t);
  fe_sub(&h, &u2, &a->x); fe_sub(&rr, &s2, &a->y);
  if (fe_is_zero(&h)) {
    if (fe_is_zero(&rr)) gej_double(r, a); else gej_set_inf(r);
    return;
  }
  fe hh, hhh, v, x3, y3, z3;
  fe_sqr(&hh, &h); fe_mul(&hhh, &hh, &h); fe_mul(&v, &a->x, &hh);
  fe_sqr(&x3, &rr); fe_sub(&x3, &x3, &hhh); fe_sub(&x3, &x3, &v); fe_sub(&x3, &x3, &v);
  fe_sub(&t, &v, &x3); fe_mul(&y3, &rr, &t); fe_mul(&t, &a->y, &hhh); fe_sub(&y3, &y3, &t);
  fe_mul(&z3, &a->z, &h);
  r->x = x3; r->y = y3; r->z = z3; r->inf = 0;
}

static void ge_set_gej(ge* r, const gej* a) {
  if (a->inf) { memset(r, 0, sizeof *r); r->inf = 1; return; }
  fe zi, zi2, zi3;
  fe_inv(&zi, &a->z); fe_sqr(&zi2, &zi); fe_mul(&zi3, &zi2, &zi);
  fe_mul(&r->x, &a->x, &zi2); fe_mul(&r->y, &a->y, &zi3); r->inf = 0;
}

/* y^2 = x^3 + 7 ; pick the root with the requested parity. Returns 0 if x is not on the curve. */
static int ge_set_xo(ge* r, const fe* x, int odd) {
  fe x2, x3, seven, y;
  fe_sqr(&x2, x); fe_mul(&x3, &x2, x);
  fe_set_int(&seven, 7); fe_add(&x3, &x3, &seven);
  if (!fe_sqrt(&y, &x3)) return 0;
  if (fe_is_odd(&y) != odd) fe_neg(&y, &y);
  r->x = *x; r->y = y; r->inf = 0;
  return 1;
}

/* ---- generator comb table: GTAB[i][j-1] = j * 2^(8 i) * G, affine ---- */
static ge (*GTAB)[255];
static pthread_once_t g_once = PTHREAD_ONCE_INIT;

static void build_gtab(void) {
  GTAB = malloc(sizeof(ge[255]) * 32);
  gej* tmp = malloc(sizeof(gej) * 32 * 255);
  gej base;
  gej_set_ge(&base, &GE_G);
  for (int i = 0; i < 32; i++) {
    ge base_aff;
    ge_set_gej(&base_aff, &base);
    gej acc;
    gej_set_ge(&acc, &base_aff);
    for (int j = 1; j <= 255; j++) {
      tmp[i * 255 + j - 1] = acc;
      gej_add_ge(&acc, &acc, &base_aff);
    }
    base = acc; /* 256 * base */
  }
  /* batch normalisation (Montgomery's trick) */
  size_t N = 32 * 255;
  fe* pre = malloc(sizeof(fe) * N);
  fe run; fe_set_int(&run, 1);
  for (size_t k = 0; k < N; k++) { pre[k] = run; fe_mul(&run, &run, &tmp[k].z); }
  fe inv; fe_inv(&inv, &run);
  for (size_t k = N; k-- > 0;) {
    fe zi, zi2, zi3;
    fe_mul(&zi, &inv, &pre[k]);
    fe_mul(&inv, &inv, &tmp[k].z);
    fe_sqr(&zi2, &zi); fe_mul(&zi3, &zi2, &zi);
    ge* o = &GTAB[k / 255][k % 255];
    fe_mul(&o->x, &tmp[k].x, &zi2); fe_mul(&o->y, &tmp[k].y, &zi3); o->inf = 0;
  }
  free(pre); free(tmp);
}
void ok_secp_init(void) { pthread_once(&g_once, build_gtab); }

static void ecmult_gen(gej* r, const sc* k) {
  ok_secp_init();
  gej_set_inf(r);
  for (int i = 0; i < 32; i++) {
    unsigned d = (unsigned)((k->n[i >> 3] >> (8 * (i & 7))) & 0xff);
    if (d) gej_add_ge(r, r, &GTAB[i][d - 1]);
  }
}

/* width-5 NAF of a 256-bit scalar; returns number of digits */
static int sc_wnaf5(int8_t out[260], const sc* k) {
  uint64_t a[5] = {k->n[0], k->n[1], k->n[2], k->n[3], 0};
  int len = 0;
  memset(out, 0, 260);
  while (a[0] | a[1] | a[2] | a[3] | a[4]) {
    int d = 0;
    if (a[0] & 1) {
      d = (int)(a[0] & 31);
      if (d >= 16) d -= 32;
      /* a -= d */
      if (d > 0) {
        uint64_t b = (uint64_t)d;
        for (int i = 0; i < 5 && b; i++) { uint64_t o = a[i]; a[i] = o - b; b = o < b; }
      } else {
        uint64_t c = (uint64_t)(-d);
        for (int i = 0; i < 5 && c; i++) { uint64_t o = a[i]; a[i] = o + c; c = a[i] < o; }
      }
    }
    out[len++] = (int8_t)d;
    for (int i = 0; i < 4; i++) a[i] = (a[i] >> 1) | (a[i + 1] << 63);
    a[4] >>= 1;
  }
  return len;
}

/* r = na * P + ng * G */
static void ecmult(gej* r, const ge* P, const sc* na, const sc* ng) {
  gej acc;
  gej_set_inf(&acc);
  if (!sc_is_zero(na) && !P->inf) {
    gej tab[8], p2, pj;
    gej_set_ge(&pj, P);
    gej_double(&p2, &pj);
    tab[0] = pj;
    for (int i = 1; i < 8; i++) gej_add(&tab[i], &tab[i - 1], &p2);
    int8_t naf[260];
    int len = sc_wnaf5(naf, na);
    for (int i = len - 1; i >= 0; i--) {
      gej_double(&acc, &acc);
      int d = naf[i];
      if (d > 0) gej_add(&acc, &acc, &tab[(d - 1) >> 1]);
      else if (d < 0) { gej t = tab[(-d - 1) >> 1]; fe_neg(&t.y, &t.y); gej_add(&acc, &acc, &t); }
    }
  }
  gej g;
  ecmult_gen(&g, ng);
  gej_add(r, &acc, &g);
}

/* ------------------------------------------------------------ BIP-340 */
static void tagged_hash(const char* tag, const uint8_t* a, size_t an, const uint8_t* b, size_t bn, const uint8_t* c, size_t cn, uint8_t out[32]) {
  uint8_t th[32];
  ok_sha256(tag, strlen(tag), th);
  ok_sha256_ctx s;
  ok_sha256_init(&s);
  ok_sha256_update(&s, th, 32);
  ok_sha256_update(&s, th, 32);
  if (an) ok_sha256_update(&s, a, an);
  if (bn) ok_sha256_update(&s, b, bn);
  if (cn) ok_sha256_update(&s, c, cn);
  ok_sha256_final(&s, out);
}

int ok_schnorr_verify(const uint8_t pk32[32], const uint8_t msg32[32], const uint8_t sig64[64]) {
  fe px, rx;
  ge P;
  /* XOnlyPublicKey::from_slice (lib.rs:582): x must be < p and lift to a curve point */
  if (!fe_set_b32_limit(&px, pk32)) return OK_SIG_PK_PARSE_ERR;
  if (!ge_set_xo(&P, &px, 0)) return OK_SIG_PK_PARSE_ERR;
  /* schnorr::Signature::from_slice (lib.rs:583) only checks the length; r >= p or s >= n
   * are rejected inside verify => "invalid", not a parse error. */
  if (!fe_set_b32_limit(&rx, sig64)) return OK_SIG_INVALID;
  sc s, e;
  int overflow;
  sc_set_b32(&s, sig64 + 32, &overflow);
  if (overflow) return OK_SIG_INVALID;
  uint8_t eh[32];
  tagged_hash("BIP0340/challenge", sig64, 32, pk32, 32, msg32, 32, eh);
  sc_set_b32(&e, eh, NULL);
  sc_neg(&e, &e);
  gej rj;
  ecmult(&rj, &P, &e, &s); /* R = s*G - e*P */
  if (rj.inf) return OK_SIG_INVALID;
  ge R;
  ge_set_gej(&R, &rj);
  if (fe_is_odd(&R.y)) return OK_SIG_INVALID;
  return fe_equal(&R.x, &rx) ? OK_SIG_VALID : OK_SIG_INVALID;
}

int ok_ecdsa_verify(const uint8_t pk33[33], const uint8_t msg32[32], const uint8_t sig64[64]) {
  fe qx;
  ge Q;
  /* PublicKey::from_slice on 33 bytes (lib.rs:618): tag 02/03, x < p, on curve */
  if (pk33[0] != 0x02 && pk33[0] != 0x03) return OK_SIG_PK_PARSE_ERR;
  if (!fe_set_b32_limit(&qx, pk33 + 1)) return OK_SIG_PK_PARSE_ERR;
  if (!ge_set_xo(&Q, &qx, pk33[0] == 0x03)) return OK_SIG_PK_PARSE_ERR;
  /* ecdsa::Signature::from_compact (lib.rs:619): r and s must each be < n */
  sc r, s, m;
  int ovr, ovs;
  sc_set_b32(&r, sig64, &ovr);
  sc_set_b32(&s, sig64 + 32, &ovs);
  if (ovr || ovs) return OK_SIG_SIG_PARSE_ERR;
  sc_set_b32(&m, msg32, NULL); /* message digest reduced mod n */
  /* verify (lib.rs:628): high S is rejected (no normalisation), r = 0 or s = 0 rejected */
  if (sc_is_high(&s)) return OK_SIG_INVALID;
  if (sc_is_zero(&r) || sc_is_zero(&s)) return OK_SIG_INVALID;
  sc sn, u1, u2;
  sc_inv(&sn, &s);
  sc_mul(&u1, &sn, &m);
  sc_mul(&u2, &sn, &r);
  gej pr;
  ecmult(&pr, &Q, &u2, &u1);
  if (pr.inf) return OK_SIG_INVALID;
  ge R;
  ge_set_gej(&R, &pr);
  /* x(R) mod n == r, with x(R) in [0,p) and r in [0,n) */
  uint64_t xr[4];
  memcpy(xr, R.x.n, sizeof xr);
  if (u256_cmp(xr, SC_N.n) >= 0) u256_sub(xr, xr, SC_N.n);
  return u256_cmp(xr, r.n) == 0 ? OK_SIG_VALID : OK_SIG_INVALID;
}

/* ------------------------------------------------------------ batch drivers */
typedef struct { const uint8_t *pk, *msg, *sig; uint8_t* st; size_t lo, hi; int ecdsa; } job_t;
static void* batch_worker(void* arg) {
  job_t* j = (job_t*)arg;
  for (size_t i = j->lo; i < j->hi; i++)
    j->st[i] = (uint8_t)(j->ecdsa ? ok_ecdsa_verify(j->pk + 33 * i, j->msg + 32 * i, j->sig + 64 * i) : ok_schnorr_verify(j->pk + 32 * i, j->msg + 32 * i, j->sig + 64 * i));
  return NULL;
}
static void run_batch(const uint8_t* pk, const uint8_t* msg, const uint8_t* sig, size_t n, uint8_t* st, int nthreads, int ecdsa) {
  ok_secp_init();
  if (nthreads < 1) nthreads = 1;
  if ((size_t)nthreads > n) nthreads = n ? (int)n : 1;
  pthread_t* th = malloc(sizeof(pthread_t) * nthreads);
  job_t* jobs = malloc(sizeof(job_t) * nthreads);
  for (int t = 0; t < nthreads; t++) {
    jobs[t] = (job_t){pk, msg, sig, st, n * t / nthreads, n * (t + 1) / nthreads, ecdsa};
    if (t + 1 < nthreads) pthread_create(&th[t], NULL, batch_worker, &jobs[t]);
  }
  batch_worker(&jobs[nthreads - 1]);
  for (int t = 0; t + 1 < nthreads; t++) pthread_join(th[t], NULL);
  free(th); free(jobs);
}
void ok_schnorr_verify_batch(const uint8_t* pk32, const uint8_t* msg32, const uint8_t* sig64, size_t n, uint8_t* status, int nthreads) { run_batch(pk32, msg32, sig64, n, status, nthreads, 0); }
void ok_ecdsa_verify_batch(const uint8_t* pk33, const uint8_t* msg32, const uint8_t* sig64, size_t n, uint8_t* status, int nthreads) { run_batch(pk33, msg32, sig64, n, status, nthreads, 1); }

/* ------------------------------------------------------------ signing (test-vector generation only) */
static int seckey_load(sc* d, const uint8_t sk[32]) {
  int ov;
  sc_set_b32(d, sk, &ov);
  return !ov && !sc_is_zero(d);
}
int ok_schnorr_pubkey(const uint8_t seckey32[32], uint8_t pk32[32]) {
  sc d;
  if (!seckey_load(&d, seckey32)) return 0;
  gej pj; ge P;
  ecmult_gen(&pj, &d); ge_set_gej(&P, &pj);
  fe_get_b32(pk32, &P.x);
  return 1;
}
int ok_schnorr_sign(const uint8_t seckey32[32], const uint8_t msg32[32], uint8_t sig64[64]) {
  sc d, k, e, s;
  if (!seckey_load(&d, seckey32)) return 0;
  gej pj; ge P;
  ecmult_gen(&pj, &d); ge_set_gej(&P, &pj);
  if (fe_is_odd(&P.y)) sc_neg(&d, &d);
  uint8_t px[32], dbytes[32], t[32], aux[32] = {0}, rnd[32], rx[32], eh[32];
  fe_get_b32(px, &P.x);
  sc_get_b32(dbytes, &d);
  tagged_hash("BIP0340/aux", aux, 32, NULL, 0, NULL, 0, t);
  for (int i = 0; i < 32; i++) t[i] ^= dbytes[i];
  tagged_hash("BIP0340/nonce", t, 32, px, 32, msg32, 32, rnd);
  sc_set_b32(&k, rnd, NULL);
  if (sc_is_zero(&k)) return 0;
  gej rj; ge R;
  ecmult_gen(&rj, &k); ge_set_gej(&R, &rj);
  if (fe_is_odd(&R.y)) sc_neg(&k, &k);
  fe_get_b32(rx, &R.x);
  tagged_hash("BIP0340/challenge", rx, 32, px, 32, msg32, 32, eh);
  sc_set_b32(&e, eh, NULL);
  sc_mul(&s, &e, &d);
  sc_add(&s, &s, &k);
  memcpy(sig64, rx, 32);
  sc_get_b32(sig64 + 32, &s);
  return 1;
}
int ok_ecdsa_pubkey(const uint8_t seckey32[32], uint8_t pk33[33]) {
  sc d;
  if (!seckey_load(&d, seckey32)) return 0;
  gej pj; ge P;
  ecmult_gen(&pj, &d); ge_set_gej(&P, &pj);
  pk33[0] = fe_is_odd(&P.y) ? 0x03 : 0x02;
  fe_get_b32(pk33 + 1, &P.x);
  return 1;
}
int ok_ecdsa_sign(const uint8_t seckey32[32], const uint8_t msg32[32], uint8_t sig64[64]) {
  sc d, k, r, s, m, ki;
  if (!seckey_load(&d, seckey32)) return 0;
  sc_set_b32(&m, msg32, NULL);
  uint8_t buf[65], kh[32];
  memcpy(buf, seckey32, 32);
  memcpy(buf + 32, msg32, 32);
  for (uint8_t ctr = 0;; ctr++) {
    buf[64] = ctr;
    ok_sha256(buf, 65, kh);
    int ov;
    sc_set_b32(&k, kh, &ov);
    if (ov || sc_is_zero(&k)) continue;
    gej rj; ge R;
    ecmult_gen(&rj, &k); ge_set_gej(&R, &rj);
    uint8_t rx[32];
    fe_get_b32(rx, &R.x);
    sc_set_b32(&r, rx, NULL);
    if (sc_is_zero(&r)) continue;
    sc_inv(&ki, &k);
    sc_mul(&s, &r, &d);
    sc_add(&s, &s, &m);
    sc_mul(&s, &s, &ki);
    if (sc_is_zero(&s)) continue;
    if (sc_is_high(&s)) sc_neg(&s, &s);
    sc_get_b32(sig64, &r);
    sc_get_b32(sig64 + 32, &s);
    return 1;
  }
}

/* ------------------------------------------------------------ probes for cross-checks */
int ok_ec_mul_xy(const uint8_t scalar32[32], const uint8_t px32[32], const uint8_t py32[32], uint8_t outx[32], uint8_t outy[32]) {
  sc k, zero = {{0, 0, 0, 0}};
  ge P, R;
  sc_set_b32(&k, scalar32, NULL);
  u256_from_be(P.x.n, px32); u256_from_be(P.y.n, py32); P.inf = 0;
  gej rj;
  ecmult(&rj, &P, &k, &zero);
  if (rj.inf) return 0;
  ge_set_gej(&R, &rj);
  fe_get_b32(outx, &R.x); fe_get_b32(outy, &R.y);
  return 1;
}
void ok_fe_mul_bytes(const uint8_t a[32], const uint8_t b[32], uint8_t out[32]) {
  fe x, y, z;
  uint64_t t[8] = {0};
  u256_from_be(t, a); fe_reduce_wide(&x, t);
  memset(t, 0, sizeof t); u256_from_be(t, b); fe_reduce_wide(&y, t);
  fe_mul(&z, &x, &y);
  fe_get_b32(out, &z);
}
void ok_sc_mul_bytes(const uint8_t a[32], const uint8_t b[32], uint8_t out[32]) {
  sc x, y, z;
  sc_set_b32(&x, a, NULL); sc_set_b32(&y, b, NULL);
  sc_mul(&z, &x, &y);
  sc_get_b32(out, &z);
}
