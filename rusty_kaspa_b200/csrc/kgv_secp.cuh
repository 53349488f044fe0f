// kgv_secp.cuh — secp256k1 group law, GLV split, window recoding and the double-scalar
// multiplication R = kP*P + kG*G used by both verifiers (BIP-340 Schnorr, ECDSA).
//
// GPU-native restructuring of what libsecp256k1's ecmult does for the reference
// (crypto/txscript/src/lib.rs:593, :628):
//   * one signature per thread, branch-uniform fixed windows (no wNAF: no per-lane divergence)
//   * P part: GLV split kP = k1 + k2*lambda (|k1|,|k2| < 2^128), signed odd 4-bit digits
//     (33 digits each, never zero), 8-entry table {1,3,..,15}*P per thread in shared memory,
//     built on an isomorphic curve so the entries are affine ("effective affine")
//   * G part: kG = lo + 2^128*hi, unsigned 16-bit windows into two 65536-entry affine tables
//     (G and 2^128*G, 8 MiB total, L2 resident), 16 mixed additions
//   * 128 shared doublings
#pragma once
#include "kgv_arith.cuh"

namespace kgv {

// ------------------------------------------------------------------------------------------
// constants
// ------------------------------------------------------------------------------------------
// group order n, little-endian limbs
#define KGV_N_LIMBS {0xD0364141u, 0xBFD25E8Cu, 0xAF48A03Bu, 0xBAAEDCE6u, 0xFFFFFFFEu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu}
// beta: cube root of unity mod p with lambda*(x,y) = (beta*x, y)
#define KGV_BETA_LIMBS {0x719501EEu, 0xC1396C28u, 0x12F58995u, 0x9CF04975u, 0xAC3434E9u, 0x6E64479Eu, 0x657C0710u, 0x7AE96A2Bu}
// GLV lattice (derived in tools/derive_constants.py): g1 = round(2^384*b2/n), g2 = round(2^384*(-b1)/n)
#define KGV_G1_LIMBS {0x45DBB031u, 0xE893209Au, 0x71E8CA7Fu, 0x3DAA8A14u, 0x9284EB15u, 0xE86C90E4u, 0xA7D46BCDu, 0x3086D221u}
#define KGV_G2_LIMBS {0x8AC47F71u, 0x1571B4AEu, 0x9DF506C6u, 0x221208ACu, 0x0ABFE4C4u, 0x6F547FA9u, 0x010E8828u, 0xE4437ED6u}
// a1 = b2 (126 bits), |b1| (128 bits), a2 (129 bits) as 5 limbs
#define KGV_A1_LIMBS {0x9284EB15u, 0xE86C90E4u, 0xA7D46BCDu, 0x3086D221u, 0u}
#define KGV_MB1_LIMBS {0x0ABFE4C3u, 0x6F547FA9u, 0x010E8828u, 0xE4437ED6u, 0u}
#define KGV_A2_LIMBS {0x9D44CFD8u, 0x57C1108Du, 0xA8E2F3F6u, 0x14CA50F7u, 1u}

#ifndef KGV_PAIRED_MUL
#define KGV_PAIRED_MUL 0   // 1: issue independent field products of the group law in pairs (fe_mul2). Measured SLOWER on B200
                           // (28.8 vs 33.7 M verifies/s: argument moves + spills outweigh the extra ILP), kept for reference.
#endif

struct gej {
  fe x, y, z;
  bool inf;
};

// Optional intermediate-value tracing (audit/debug kernels and the host unit tests use it to
// compare device and host executions stage by stage); NoTrace compiles to nothing.
struct NoTrace {
  KGV_HD void operator()(int /*stage*/, const uint32_t* /*words*/, int /*n*/) const {}
};

// ------------------------------------------------------------------------------------------
// scalars (mod n): only what verification needs
// ------------------------------------------------------------------------------------------
// a >= n ?
KGV_HD bool sc_ge_n(const uint32_t* a) {
  const uint32_t n[8] = KGV_N_LIMBS;
  return !lt8(a, n);
}
// a in [0, 2^256) -> a mod n (2^256 < 2n: one conditional subtraction)
KGV_HD void sc_reduce_once(uint32_t* a) {
  const uint32_t n[8] = KGV_N_LIMBS;
  uint32_t t[8];
  uint32_t bo = sub8(t, a, n);
  if (!bo) {
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = t[i];
  }
}
KGV_HD bool is_zero8(const uint32_t* a) { return (a[0] | a[1] | a[2] | a[3] | a[4] | a[5] | a[6] | a[7]) == 0; }
// r = -a mod n, for a in [0,n)
KGV_HD void sc_neg(uint32_t* r, const uint32_t* a) {
  const uint32_t n[8] = KGV_N_LIMBS;
  if (is_zero8(a)) {
#pragma unroll
    for (int i = 0; i < 8; i++) r[i] = 0;
    return;
  }
  (void)sub8(r, n, a);
}

// ---- full scalar multiplication / inversion mod n (ECDSA only) ----
// acc[0..NA) += x[0..NX) * (2^256 - n); carries propagate to the top limb
template <int NA, int NX>
KGV_HD void sc_fold(uint32_t* acc, const uint32_t* x) {
  const uint32_t nc[5] = {0x2FC9BEBFu, 0x402DA173u, 0x50B75FC4u, 0x45512319u, 1u};
#pragma unroll
  for (int i = 0; i < NX; i++) {
    uint64_t c = 0;
#pragma unroll
    for (int j = 0; j < 5; j++) {
      if (i + j < NA) {
        uint64_t t = (uint64_t)x[i] * nc[j] + acc[i + j] + c;
        acc[i + j] = (uint32_t)t;
        c = t >> 32;
      }
    }
#pragma unroll
    for (int k = i + 5; k < NA; k++) {
      c += acc[k];
      acc[k] = (uint32_t)c;
      c >>= 32;
    }
  }
}
// r = t mod n for a 512-bit t
KGV_HD void sc_reduce512(uint32_t* r, const uint32_t* t) {
  uint32_t a[13], b[9], c[9];
#pragma unroll
  for (int i = 0; i < 13; i++) a[i] = i < 8 ? t[i] : 0u;
  sc_fold<13, 8>(a, t + 8);          // < 2^386
#pragma unroll
  for (int i = 0; i < 9; i++) b[i] = i < 8 ? a[i] : 0u;
  sc_fold<9, 5>(b, a + 8);           // < 2^260
#pragma unroll
  for (int i = 0; i < 9; i++) c[i] = i < 8 ? b[i] : 0u;
  sc_fold<9, 1>(c, b + 8);           // < 2^256 + 2^133
#pragma unroll
  for (int i = 0; i < 8; i++) r[i] = c[i];
  uint32_t top = c[8];               // 0 or 1; if 1 the low part is tiny
  {
    uint32_t one[1] = {top};
    uint32_t d[8];
#pragma unroll
    for (int i = 0; i < 8; i++) d[i] = r[i];
    sc_fold<8, 1>(d, one);
#pragma unroll
    for (int i = 0; i < 8; i++) r[i] = d[i];
  }
  sc_reduce_once(r);
}
// As for fe_mul / fe_sqr, the device versions are real functions (operands by value, in registers).
// KGV_SC_ONE_BLOCK (default): ALL scalar arithmetic of the ECDSA kernel goes through ONE non-inlined function, r = a^(2^k) * (b or 1), with a
// single copy of the 8x8 product and of the reduction (squarings use the general product: +28 multiplies each, 0.4 % more instructions per
// verification).  Three separate blocks (multiply, square, run-of-squarings: ~19 KB of SASS next to the 25 KB of the point arithmetic) pushed
// the kernel's hot code out of the instruction cache: ncu showed `no_instruction` stalls at 1.88 per issued instruction against 0.75 in the
// Schnorr kernel.
#ifndef KGV_SC_ONE_BLOCK
#define KGV_SC_ONE_BLOCK 1
#endif
struct sc8 { uint32_t v[8]; };
#if defined(__CUDACC__) && KGV_NOINLINE_MUL && KGV_SC_ONE_BLOCK
static __device__ __noinline__ sc8 sc_pow2k_mul_call(sc8 a, int k, int with_mul, sc8 b) {
  const int n = k + with_mul;
#pragma unroll 1
  for (int i = 0; i < n; i++) {
    sc8 y;
#pragma unroll
    for (int w = 0; w < 8; w++) y.v[w] = i < k ? a.v[w] : b.v[w];
    uint32_t t[16];
    mul_wide(t, a.v, y.v);
    sc_reduce512(a.v, t);
  }
  return a;
}
KGV_HD void sc_mul(uint32_t* r, const uint32_t* a, const uint32_t* b) {
  sc8 x, y;
#pragma unroll
  for (int i = 0; i < 8; i++) { x.v[i] = a[i]; y.v[i] = b[i]; }
  sc8 z = sc_pow2k_mul_call(x, 0, 1, y);
#pragma unroll
  for (int i = 0; i < 8; i++) r[i] = z.v[i];
}
KGV_HD void sc_sqr(uint32_t* r, const uint32_t* a) {
  sc8 x;
#pragma unroll
  for (int i = 0; i < 8; i++) x.v[i] = a[i];
  sc8 z = sc_pow2k_mul_call(x, 1, 0, x);
#pragma unroll
  for (int i = 0; i < 8; i++) r[i] = z.v[i];
}
// r = r^(2^k) * m
KGV_HD void sc_sqr_n_mul(uint32_t* r, int k, const uint32_t* m) {
  sc8 x, y;
#pragma unroll
  for (int i = 0; i < 8; i++) { x.v[i] = r[i]; y.v[i] = m[i]; }
  sc8 z = sc_pow2k_mul_call(x, k, 1, y);
#pragma unroll
  for (int i = 0; i < 8; i++) r[i] = z.v[i];
}
#elif defined(__CUDACC__) && KGV_NOINLINE_MUL
static __device__ __noinline__ sc8 sc_mul_call(sc8 a, sc8 b) {
  sc8 r;
  uint32_t t[16];
  mul_wide(t, a.v, b.v);
  sc_reduce512(r.v, t);
  return r;
}
static __device__ __noinline__ sc8 sc_sqr_call(sc8 a) {
  sc8 r;
  uint32_t t[16];
  sqr_wide(t, a.v);
  sc_reduce512(r.v, t);
  return r;
}
KGV_HD void sc_mul(uint32_t* r, const uint32_t* a, const uint32_t* b) {
  sc8 x, y;
#pragma unroll
  for (int i = 0; i < 8; i++) { x.v[i] = a[i]; y.v[i] = b[i]; }
  sc8 z = sc_mul_call(x, y);
#pragma unroll
  for (int i = 0; i < 8; i++) r[i] = z.v[i];
}
KGV_HD void sc_sqr(uint32_t* r, const uint32_t* a) {
  sc8 x;
#pragma unroll
  for (int i = 0; i < 8; i++) x.v[i] = a[i];
  sc8 z = sc_sqr_call(x);
#pragma unroll
  for (int i = 0; i < 8; i++) r[i] = z.v[i];
}
#else
KGV_HD void sc_mul(uint32_t* r, const uint32_t* a, const uint32_t* b) {
  uint32_t t[16];
  mul_wide(t, a, b);
  sc_reduce512(r, t);
}
KGV_HD void sc_sqr(uint32_t* r, const uint32_t* a) {
  uint32_t t[16];
  sqr_wide(t, a);
  sc_reduce512(r, t);
}
#endif
// r = a^(n-2) mod n (a != 0).  The exponent is public and identical in every lane: no divergence.  Addition chain: the 127 leading one bits
// of n-2 through x_k = a^(2^k - 1) (k = 2, 3, 6, 8, 14, 28, 56, 112, 126: the ladder libsecp256k1's scalar inverse uses), the remaining 129 bits
// by a sliding window over the odd powers a, a^3, a^5, a^7 (schedule derived and checked against pow(a, n-2, n) by
// tools/derive_sc_inv_chain.py; tests/test_hostsim.py runs this very function on the host): 255 squarings + 44 multiplications instead of the 255 + 191
// of plain square-and-multiply (ECDSA shares one inversion among KGV_ITEMS signatures; it was 11 % of an ECDSA verification).
#if defined(__CUDACC__) && KGV_NOINLINE_MUL && KGV_SC_ONE_BLOCK
// (sc_sqr_n_mul above)
#else
#if defined(__CUDACC__) && KGV_NOINLINE_MUL && KGV_SC_SQRN_CALL
static __device__ __noinline__ sc8 sc_sqr_n_call(sc8 a, int n) {  // one call per run of squarings (see fe_sqr_n)
#pragma unroll 1
  for (int i = 0; i < n; i++) {
    uint32_t t[16];
    sqr_wide(t, a.v);
    sc_reduce512(a.v, t);
  }
  return a;
}
KGV_HD void sc_sqr_n(uint32_t* r, int n) {
  sc8 x;
#pragma unroll
  for (int i = 0; i < 8; i++) x.v[i] = r[i];
  x = sc_sqr_n_call(x, n);
#pragma unroll
  for (int i = 0; i < 8; i++) r[i] = x.v[i];
}
#else
KGV_HD void sc_sqr_n(uint32_t* r, int n) {
  for (int i = 0; i < n; i++) sc_sqr(r, r);
}
#endif
KGV_HD void sc_sqr_n_mul(uint32_t* r, int k, const uint32_t* m) { sc_sqr_n(r, k); sc_mul(r, r, m); }
#endif
KGV_HD void sc_inv(uint32_t* r, const uint32_t* a) {
  uint32_t a1[8], a3[8], a5[8], a7[8], x6[8], x14[8], t[8], u[8];
#pragma unroll
  for (int i = 0; i < 8; i++) a1[i] = a[i];
  sc_sqr(u, a1);                       // a^2
  sc_mul(a3, u, a1);                   // x2 = a^3
  sc_mul(a5, a3, u);                   // a^5
  sc_sqr(t, a3); sc_mul(a7, t, a1);    // x3 = a^7
#pragma unroll
  for (int i = 0; i < 8; i++) t[i] = a7[i];
  sc_sqr_n_mul(t, 3, a7);                                              // x6
#pragma unroll
  for (int i = 0; i < 8; i++) x6[i] = t[i];
  sc_sqr_n_mul(t, 2, a3);                                              // x8
  sc_sqr_n_mul(t, 6, x6);                                              // x14
#pragma unroll
  for (int i = 0; i < 8; i++) x14[i] = t[i];
  sc_sqr_n_mul(t, 14, x14);                                            // x28
#pragma unroll
  for (int i = 0; i < 8; i++) u[i] = t[i];
  sc_sqr_n_mul(t, 28, u);                                              // x56
#pragma unroll
  for (int i = 0; i < 8; i++) u[i] = t[i];
  sc_sqr_n_mul(t, 56, u);                                              // x112
  sc_sqr_n_mul(t, 14, x14);                                            // x126
  sc_sqr_n_mul(t, 1, a1);                                              // the 127 leading ones
#define SC_STEP(k, x) sc_sqr_n_mul(t, k, x);
  SC_STEP(4, a5) SC_STEP(2, a3) SC_STEP(4, a5) SC_STEP(4, a5) SC_STEP(2, a3) SC_STEP(3, a3)
  SC_STEP(4, a7) SC_STEP(5, a7) SC_STEP(4, a3) SC_STEP(4, a5) SC_STEP(4, a7) SC_STEP(3, a5)
  SC_STEP(3, a1) SC_STEP(6, a5) SC_STEP(10, a7) SC_STEP(4, a7) SC_STEP(4, a7) SC_STEP(3, a7)
  SC_STEP(2, a3) SC_STEP(2, a1) SC_STEP(3, a1) SC_STEP(5, a5) SC_STEP(3, a7) SC_STEP(2, a1)
  SC_STEP(5, a3) SC_STEP(4, a3) SC_STEP(2, a1) SC_STEP(8, a3) SC_STEP(3, a3) SC_STEP(3, a1)
  SC_STEP(6, a1) SC_STEP(5, a7) SC_STEP(3, a7)
#undef SC_STEP
#pragma unroll
  for (int i = 0; i < 8; i++) r[i] = t[i];
}
// a > (n-1)/2 ?
KGV_HD bool sc_is_high(const uint32_t* a) {
  const uint32_t hn[8] = {0x681B20A0u, 0xDFE92F46u, 0x57A4501Du, 0x5D576E73u, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0x7FFFFFFFu};
  return lt8(hn, a);
}

// r[0..4] = (a[0..4] * b[0..4]) mod 2^160   (15 partial products)
KGV_HD void mul_trunc5(uint32_t* r, const uint32_t* a, const uint32_t* b) {
#pragma unroll
  for (int i = 0; i < 5; i++) r[i] = 0;
#pragma unroll
  for (int i = 0; i < 5; i++) {
    uint64_t c = 0;
#pragma unroll
    for (int j = 0; j + i < 5; j++) {
      uint64_t t = (uint64_t)a[i] * b[j] + r[i + j] + c;
      r[i + j] = (uint32_t)t;
      c = t >> 32;
    }
  }
}

// 160-bit two's complement helpers (5 limbs)
KGV_HD void sub5(uint32_t* r, const uint32_t* a, const uint32_t* b) {
  uint64_t br = 0;
#pragma unroll
  for (int i = 0; i < 5; i++) {
    uint64_t t = (uint64_t)a[i] - b[i] - br;
    r[i] = (uint32_t)t;
    br = (t >> 32) & 1;
  }
}
KGV_HD void neg5(uint32_t* r) {
  uint64_t c = 1;
#pragma unroll
  for (int i = 0; i < 5; i++) {
    c += (uint32_t)~r[i];
    r[i] = (uint32_t)c;
    c >>= 32;
  }
}

// GLV split: k (mod n, 8 limbs) -> |k1|, |k2| (5 limbs each, < 2^128 in practice) and their signs,
// with k == s1*|k1| + s2*|k2|*lambda (mod n).  Pure integer arithmetic (no modular reduction):
//   c1 = round(k*g1 / 2^384), c2 = round(k*g2 / 2^384)
//   k1 = k - c1*a1 - c2*a2,  k2 = c1*|b1| - c2*b2      (both tiny, evaluated mod 2^160)
KGV_HD void glv_split(uint32_t* k1, bool& neg1, uint32_t* k2, bool& neg2, const uint32_t* k) {
  const uint32_t g1[8] = KGV_G1_LIMBS, g2[8] = KGV_G2_LIMBS;
  const uint32_t a1[5] = KGV_A1_LIMBS, mb1[5] = KGV_MB1_LIMBS, a2[5] = KGV_A2_LIMBS;
  uint32_t t[16], c1[5], c2[5];
  mul_wide(t, k, g1);
  {
    uint64_t c = (t[11] >> 31);
#pragma unroll
    for (int i = 0; i < 4; i++) { c += t[12 + i]; c1[i] = (uint32_t)c; c >>= 32; }
    c1[4] = (uint32_t)c;
  }
  mul_wide(t, k, g2);
  {
    uint64_t c = (t[11] >> 31);
#pragma unroll
    for (int i = 0; i < 4; i++) { c += t[12 + i]; c2[i] = (uint32_t)c; c >>= 32; }
    c2[4] = (uint32_t)c;
  }
  uint32_t p1[5], p2[5];
  // k1 = k - c1*a1 - c2*a2  (mod 2^160)
  mul_trunc5(p1, c1, a1);
  mul_trunc5(p2, c2, a2);
  sub5(k1, k, p1);
  sub5(k1, k1, p2);
  // k2 = c1*|b1| - c2*b2  (b2 == a1)
  mul_trunc5(p1, c1, mb1);
  mul_trunc5(p2, c2, a1);
  sub5(k2, p1, p2);
  neg1 = (k1[4] >> 31) != 0;
  if (neg1) neg5(k1);
  neg2 = (k2[4] >> 31) != 0;
  if (neg2) neg5(k2);
}

// Signed odd-digit recoding of a magnitude m < 2^131 (5 limbs).
// If m is even it is replaced by m+1 and `fix` is set (caller subtracts one table base point).
// Then m = sum_{i=0}^{32} d_i 16^i with every d_i odd in {+-1,..,+-15}:
//   h = (m >> 1) | 2^131 ; v_i = (h >> 4i) & 15 ; d_i = 2 v_i - 15.
KGV_HD void recode_signed_odd(uint32_t* h, bool& fix, const uint32_t* m) {
  uint32_t t[5];
  fix = (m[0] & 1u) == 0;
  uint64_t c = fix ? 1 : 0;
#pragma unroll
  for (int i = 0; i < 5; i++) { c += m[i]; t[i] = (uint32_t)c; c >>= 32; }
#pragma unroll
  for (int i = 0; i < 4; i++) h[i] = (t[i] >> 1) | (t[i + 1] << 31);
  h[4] = (t[4] >> 1) | (1u << 3);  // bit 131 = bit 3 of limb 4
}
// digit i (0..32) of a recoded scalar: table index 0..7 ((|d|-1)/2) and sign
KGV_HD void recoded_digit(const uint32_t* h, int i, uint32_t& idx, bool& neg) {
  uint32_t v = (h[i >> 3] >> ((i & 7) * 4)) & 15u;
  neg = (v & 8u) == 0;
  idx = neg ? (~v & 7u) : (v & 7u);
}

// ------------------------------------------------------------------------------------------
// group law, Jacobian coordinates on y^2 = x^3 + b (a = 0; b never appears in the formulas,
// so the same code runs on the isomorphic curves used for the per-thread tables)
// ------------------------------------------------------------------------------------------
#ifndef KGV_INLINE_MUL_IN_POINT
#define KGV_INLINE_MUL_IN_POINT 0  // 1: the field products inside gej_double / gej_add_ge are inlined into those (non-inlined) functions
#endif
#if defined(__CUDACC__) && KGV_INLINE_MUL_IN_POINT
#define FE_PMUL fe_mul_inl
#define FE_PSQR fe_sqr_inl
#else
#define FE_PMUL fe_mul
#define FE_PSQR fe_sqr
#endif
KGV_HD void gej_double_body(gej& r) {
  if (r.inf) return;
  fe A, B, C, D, E, F, t, yz;
#if KGV_PAIRED_MUL
  fe_sqr2(A, r.x, B, r.y);       // A = X^2, B = Y^2
  fe_add(t, r.x, B);
  fe_sqr2(C, B, t, t);           // C = Y^4, t = (X+B)^2
  fe_sub(t, t, A);
  fe_sub(t, t, C);
  fe_dbl(D, t);                  // D = 4 X Y^2
  fe_mul3(E, A);                 // E = 3 X^2
  fe_mulsqr(yz, r.y, r.z, F, E); // Y*Z, E^2
  fe_dbl(r.z, yz);               // Z3 = 2 Y Z
  fe_sub(t, F, D);
  fe_sub(r.x, t, D);             // X3 = E^2 - 2D
  fe_sub(t, D, r.x);
  FE_PMUL(t, E, t);
  fe_mul8(C, C);
  fe_sub(r.y, t, C);             // Y3 = E (D - X3) - 8 Y^4
#else
  FE_PSQR(A, r.x);
  FE_PSQR(B, r.y);
  FE_PSQR(C, B);
  fe_add(t, r.x, B);
  FE_PSQR(t, t);
  fe_sub(t, t, A);
  fe_sub(t, t, C);
  fe_dbl(D, t);        // D = 2((X+B)^2 - A - C) = 4 X Y^2
  fe_mul3(E, A);       // E = 3 X^2
  FE_PMUL(r.z, r.y, r.z);
  fe_dbl(r.z, r.z);    // Z3 = 2 Y Z
  FE_PSQR(t, E);
  fe_sub(t, t, D);
  fe_sub(r.x, t, D);   // X3 = E^2 - 2D
  fe_sub(t, D, r.x);
  FE_PMUL(t, E, t);
  fe_mul8(C, C);
  fe_sub(r.y, t, C);   // Y3 = E (D - X3) - 8 Y^4
  (void)F; (void)yz;
#endif
}

#ifndef KGV_NOINLINE_POINT
#define KGV_NOINLINE_POINT 1
#endif
#if defined(__CUDACC__) && KGV_NOINLINE_POINT
static __device__ __noinline__ gej gej_double_call(gej r) { gej_double_body(r); return r; }
KGV_HD void gej_double(gej& r) { r = gej_double_call(r); }
#else
KGV_HD void gej_double(gej& r) { gej_double_body(r); }
#endif

// r += (bx,by) with the addend affine and never the point at infinity.  Handles r = inf,
// r == addend (doubling) and r == -addend (result infinity).  If hout != nullptr it receives
// the factor by which Z was multiplied (H), used by the table builder.
KGV_HD void gej_add_ge_body(gej& r, const fe& bx, const fe& by, fe* hout) {
  if (r.inf) {
    r.x = bx;
    r.y = by;
    fe_set_u32(r.z, 1);
    r.inf = false;
    if (hout) fe_set_u32(*hout, 1);
    return;
  }
  fe z1z1, u2, s2, h, rr, t;
  FE_PSQR(z1z1, r.z);
#if KGV_PAIRED_MUL
  fe_mul2(u2, bx, z1z1, t, r.z, z1z1);
#else
  FE_PMUL(u2, bx, z1z1);
  FE_PMUL(t, r.z, z1z1);
#endif
  FE_PMUL(s2, by, t);
  fe_sub(h, u2, r.x);
  fe_sub(rr, s2, r.y);
  if (fe_is_zero(h)) {
    if (hout) fe_set_u32(*hout, 1);
    if (fe_is_zero(rr)) {
#if KGV_INLINE_MUL_IN_POINT
      gej_double(r);                 // rare path: through the (non-inlined) doubling, not a second inlined copy of five products
#else
      gej_double_body(r);            // inlined: a CALL here would make gej_add_ge_call a non-leaf function (return-address / register saves
                                     // on EVERY addition: +1.6 % instructions, -2.5 % throughput, measured in round 2)
#endif
      if (hout) fe_dbl(*hout, r.y);  // not used by the table builder (cannot happen there)
    } else {
      r.inf = true;
    }
    return;
  }
  if (hout) *hout = h;
  fe hh, hhh, v;
#if KGV_PAIRED_MUL
  fe y1h3;
  fe_mulsqr(r.z, r.z, h, hh, h);          // Z3 = Z1*H, HH = H^2
  fe_mul2(hhh, hh, h, v, r.x, hh);        // H^3, V = X1*HH
  fe_mulsqr(y1h3, r.y, hhh, t, rr);       // Y1*H^3, R^2
  fe_sub(t, t, hhh);
  fe_sub(t, t, v);
  fe_sub(r.x, t, v);                      // X3 = R^2 - H^3 - 2V
  fe_sub(t, v, r.x);
  FE_PMUL(t, rr, t);
  fe_sub(r.y, t, y1h3);                   // Y3 = R (V - X3) - Y1 H^3
#else
  FE_PSQR(hh, h);
  FE_PMUL(hhh, hh, h);
  FE_PMUL(v, r.x, hh);
  FE_PMUL(r.z, r.z, h);
  FE_PSQR(t, rr);
  fe_sub(t, t, hhh);
  fe_sub(t, t, v);
  fe_sub(r.x, t, v);      // X3 = R^2 - H^3 - 2V
  fe_sub(t, v, r.x);
  FE_PMUL(t, rr, t);
  FE_PMUL(hhh, r.y, hhh);
  fe_sub(r.y, t, hhh);    // Y3 = R (V - X3) - Y1 H^3
#endif
}

// KGV_ADD_ONE_BLOCK (default): ONE copy of the mixed addition in the kernel.  The table builder needs the addition's H value and used to get
// it from an INLINED body - seven unrolled copies, ~160 KB of straight-line code that every signature streamed through once, evicting the
// ladder's hot code from the instruction cache.  Now the one non-inlined function always returns H as well (8 more
// registers by value, ignored by the ladder).
#ifndef KGV_ADD_ONE_BLOCK
#define KGV_ADD_ONE_BLOCK 1
#endif
#if defined(__CUDACC__) && KGV_NOINLINE_POINT && KGV_ADD_ONE_BLOCK
struct gej_h { gej r; fe h; };
static __device__ __noinline__ gej_h gej_add_ge_call(gej r, fe bx, fe by) { gej_h o; gej_add_ge_body(r, bx, by, &o.h); o.r = r; return o; }
KGV_HD void gej_add_ge(gej& r, const fe& bx, const fe& by, fe* hout = nullptr) {
  gej_h o = gej_add_ge_call(r, bx, by);
  r = o.r;
  if (hout) *hout = o.h;
}
#elif defined(__CUDACC__) && KGV_NOINLINE_POINT
static __device__ __noinline__ gej gej_add_ge_call(gej r, fe bx, fe by) { gej_add_ge_body(r, bx, by, nullptr); return r; }
#if KGV_INLINE_MUL_IN_POINT
struct gej_h { gej r; fe h; };
static __device__ __noinline__ gej_h gej_add_ge_h_call(gej r, fe bx, fe by) { gej_h o; gej_add_ge_body(r, bx, by, &o.h); o.r = r; return o; }
#endif
KGV_HD void gej_add_ge(gej& r, const fe& bx, const fe& by, fe* hout = nullptr) {
#if KGV_INLINE_MUL_IN_POINT
  if (hout) { gej_h o = gej_add_ge_h_call(r, bx, by); r = o.r; *hout = o.h; }
#else
  if (hout) gej_add_ge_body(r, bx, by, hout);
#endif
  else r = gej_add_ge_call(r, bx, by);
}
#else
KGV_HD void gej_add_ge(gej& r, const fe& bx, const fe& by, fe* hout = nullptr) { gej_add_ge_body(r, bx, by, hout); }
#endif

// y^2 = x^3 + 7: solve for y with the requested parity.  false if x is not on the curve.
// x must be canonical (< p).
KGV_HD bool ge_lift_x(fe& y, const fe& x, bool odd) {
  fe t, c;
  fe_sqr(t, x);
  fe_mul(t, t, x);
  fe_set_u32(c, 7);
  fe_add(c, t, c);
  if (!fe_sqrt(y, c)) return false;
  fe_normalize(y);
  if (((y.v[0] & 1u) != 0) != odd) {
    fe_neg(y, y);
    fe_normalize(y);
  }
  return true;
}

// ------------------------------------------------------------------------------------------
// per-thread table of odd multiples {1,3,...,15} * P, "effective affine"
// ------------------------------------------------------------------------------------------
// Tab is an accessor with  void put(int entry, int word, uint32_t v)  and  uint32_t get(int entry, int word)
// (entry 0..7, word 0..15: x limbs then y limbs).
//
// After the call, entry j holds the affine coordinates of (2j+1)*P on the isomorphic curve
// E' : y^2 = x^3 + 7*zs^6, where zs (returned) is such that a Jacobian point (X,Y,Z) on E'
// corresponds to (X, Y, Z*zs) on secp256k1.  lambda*(entry) = (beta*x, y) also holds on E'.
template <class Tab>
KGV_HD void build_odd_table(Tab& tab, fe& zs, const fe& px, const fe& py) {
  // D = 2P (from affine input)
  gej d;
  d.x = px; d.y = py; fe_set_u32(d.z, 1); d.inf = false;
  gej_double(d);
  // map P onto the curve where D is affine: (x * Zd^2, y * Zd^3)
  fe zd2, zd3;
  fe_sqr(zd2, d.z);
  fe_mul(zd3, zd2, d.z);
  gej t;
  fe_mul(t.x, px, zd2);
  fe_mul(t.y, py, zd3);
  fe_set_u32(t.z, 1);
  t.inf = false;
  fe H[7];
#pragma unroll
  for (int w = 0; w < 8; w++) { tab.put(0, w, t.x.v[w]); tab.put(0, 8 + w, t.y.v[w]); }
#pragma unroll
  for (int j = 1; j < 8; j++) {
    gej_add_ge(t, d.x, d.y, &H[j - 1]);
#pragma unroll
    for (int w = 0; w < 8; w++) { tab.put(j, w, t.x.v[w]); tab.put(j, 8 + w, t.y.v[w]); }
  }
  // bring every entry to the Z of the last one: entry j *= (Z_7/Z_j)^{2,3}, Z_7/Z_j = H_{j+1}...H_7
  fe acc = H[6];
#pragma unroll
  for (int j = 6; j >= 0; j--) {
    if (j < 6) fe_mul(acc, acc, H[j]);
    fe a2, a3, ex, ey;
    fe_sqr(a2, acc);
    fe_mul(a3, a2, acc);
#pragma unroll
    for (int w = 0; w < 8; w++) { ex.v[w] = tab.get(j, w); ey.v[w] = tab.get(j, 8 + w); }
    fe_mul(ex, ex, a2);
    fe_mul(ey, ey, a3);
#pragma unroll
    for (int w = 0; w < 8; w++) { tab.put(j, w, ex.v[w]); tab.put(j, 8 + w, ey.v[w]); }
  }
  // total Z scale back to secp256k1: Z_7 (on D's curve) * Zd
  fe_mul(zs, t.z, d.z);
}

// ------------------------------------------------------------------------------------------
// R = kP * P + kG * G     (result on the isomorphic curve; true Z = R.z * zs)
// ------------------------------------------------------------------------------------------
// gtab: [2][65536][16] u32 — affine (x limbs, y limbs) of v*G and v*2^128*G; entry 0 unused.
// GLoad is a functor  void operator()(fe& x, fe& y, const uint32_t* entry)  (vectorised loads on device).
template <class Tab, class GLoad, class Trace = NoTrace>
KGV_HD void ecmult_double(gej& R, fe& zs, const fe& px, const fe& py, const uint32_t* kP, const uint32_t* kG, Tab& tab,
                          const uint32_t* gtab, GLoad gload, Trace trace = Trace()) {
  const fe beta = {KGV_BETA_LIMBS};
  uint32_t m1[5], m2[5], h1[5], h2[5];
  bool neg1, neg2, fix1, fix2;
  glv_split(m1, neg1, m2, neg2, kP);
  recode_signed_odd(h1, fix1, m1);
  recode_signed_odd(h2, fix2, m2);
  trace(10, m1, 5); trace(11, m2, 5);
  { uint32_t f[4] = {neg1, neg2, fix1, fix2}; trace(12, f, 4); }
  build_odd_table(tab, zs, px, py);
  trace(13, zs.v, 8);
  { uint32_t e0[16]; for (int w = 0; w < 16; w++) e0[w] = tab.get(7, w); trace(14, e0, 16); }
  fe zs2, zs3;
  fe_sqr(zs2, zs);
  fe_mul(zs3, zs2, zs);

  R.inf = true;
  fe_set_zero(R.x); fe_set_zero(R.y); fe_set_zero(R.z);
  for (int i = 32; i >= 0; i--) {
    if (i != 32) {
      gej_double(R); gej_double(R); gej_double(R); gej_double(R);
    }
    uint32_t idx; bool dn;
    fe ex, ey;
    // k1 digit on P
    recoded_digit(h1, i, idx, dn);
#pragma unroll
    for (int w = 0; w < 8; w++) { ex.v[w] = tab.get(idx, w); ey.v[w] = tab.get(idx, 8 + w); }
    if (dn != neg1) fe_neg(ey, ey);
    gej_add_ge(R, ex, ey);
    // k2 digit on lambda*P = (beta*x, y)
    recoded_digit(h2, i, idx, dn);
#pragma unroll
    for (int w = 0; w < 8; w++) { ex.v[w] = tab.get(idx, w); ey.v[w] = tab.get(idx, 8 + w); }
    fe_mul(ex, ex, beta);
    if (dn != neg2) fe_neg(ey, ey);
    gej_add_ge(R, ex, ey);
    // generator part: 16-bit windows every 4th step
    if ((i & 3) == 0 && i < 32) {
      int w16 = i >> 2;  // 0..7
      uint32_t dlo = (kG[w16 >> 1] >> ((w16 & 1) * 16)) & 0xFFFFu;
      uint32_t dhi = (kG[4 + (w16 >> 1)] >> ((w16 & 1) * 16)) & 0xFFFFu;
      if (dlo) {
        gload(ex, ey, gtab + (size_t)dlo * 16);
        fe_mul(ex, ex, zs2);
        fe_mul(ey, ey, zs3);
        gej_add_ge(R, ex, ey);
      }
      if (dhi) {
        gload(ex, ey, gtab + ((size_t)65536 + dhi) * 16);
        fe_mul(ex, ex, zs2);
        fe_mul(ey, ey, zs3);
        gej_add_ge(R, ex, ey);
      }
    }
  }
  trace(15, R.x.v, 8); trace(16, R.y.v, 8); trace(17, R.z.v, 8);
  // parity corrections: m was replaced by m+1 => subtract one base point
  if (fix1) {
    fe ex, ey;
#pragma unroll
    for (int w = 0; w < 8; w++) { ex.v[w] = tab.get(0, w); ey.v[w] = tab.get(0, 8 + w); }
    if (!neg1) fe_neg(ey, ey);
    gej_add_ge(R, ex, ey);
  }
  if (fix2) {
    fe ex, ey;
#pragma unroll
    for (int w = 0; w < 8; w++) { ex.v[w] = tab.get(0, w); ey.v[w] = tab.get(0, 8 + w); }
    fe_mul(ex, ex, beta);
    if (!neg2) fe_neg(ey, ey);
    gej_add_ge(R, ex, ey);
  }
}

}  // namespace kgv
