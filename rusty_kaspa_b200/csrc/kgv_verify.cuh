// kgv_verify.cuh — per-signature verification cores (one signature per thread).
//
// Tri-state verdicts mirror what the reference's script engine can observe
// (crypto/txscript/src/lib.rs:574-643):
//   0 invalid            sig.verify() returned Err            -> Ok(false)
//   1 valid              sig.verify() returned Ok             -> Ok(true)
//   2 pubkey parse error XOnlyPublicKey/PublicKey::from_slice -> Err(InvalidSignature), script aborts
//   3 sig parse error    ecdsa::Signature::from_compact       -> Err(InvalidSignature), script aborts
#pragma once
#include "kgv_secp.cuh"
#include "kgv_sha256.cuh"

namespace kgv {

enum : uint8_t { KGV_ST_INVALID = 0, KGV_ST_VALID = 1, KGV_ST_PK_PARSE = 2, KGV_ST_SIG_PARSE = 3 };

KGV_HD bool fe_words_lt_p(const uint32_t* v) {
  const uint32_t p[8] = {KGV_P0, KGV_P1, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu};
  return lt8(v, p);
}

#define KGV_ST_PENDING 0xFFu  // phase 1 passed: the verdict needs the (batched) inversion of zt

// BIP-340 verification, phase 1: everything up to the projective result R = s*G - e*P.
// pkw/mw: 8 big-endian words, sigw: 16 big-endian words (r || s).
// Returns a final verdict, or KGV_ST_PENDING with (X, Y, zt = true Z, rx) filled in.
template <class Tab, class GLoad, class Trace = NoTrace>
KGV_HD uint8_t schnorr_phase1(fe& X, fe& Y, fe& zt, fe& rx, const uint32_t* pkw, const uint32_t* mw, const uint32_t* sigw, Tab& tab,
                              const uint32_t* gtab, GLoad gload, Trace trace = Trace()) {
  fe px, py;
  limbs_from_be_words(px.v, pkw);
  if (!fe_words_lt_p(px.v)) return KGV_ST_PK_PARSE;      // x >= p
  trace(1, px.v, 8);
  if (!ge_lift_x(py, px, false)) return KGV_ST_PK_PARSE;  // not on the curve
  trace(2, py.v, 8);
  limbs_from_be_words(rx.v, sigw);
  if (!fe_words_lt_p(rx.v)) return KGV_ST_INVALID;        // r >= p
  uint32_t s[8], e[8], k[8], ew[8];
  limbs_from_be_words(s, sigw + 8);
  if (sc_ge_n(s)) return KGV_ST_INVALID;                  // s >= n
  bip340_challenge(ew, sigw, pkw, mw);
#pragma unroll
  for (int i = 0; i < 8; i++) e[7 - i] = ew[i];
  sc_reduce_once(e);
  sc_neg(k, e);                                           // R = s*G - e*P
  trace(3, e, 8); trace(4, k, 8); trace(5, s, 8);
  gej R;
  fe zs;
  ecmult_double(R, zs, px, py, k, s, tab, gtab, gload, trace);
  { uint32_t f[1] = {R.inf}; trace(18, f, 1); }
  if (R.inf) return KGV_ST_INVALID;
  trace(19, R.x.v, 8); trace(20, R.y.v, 8); trace(21, R.z.v, 8);
  X = R.x;
  Y = R.y;
  fe_mul(zt, R.z, zs);
  return KGV_ST_PENDING;
}
// phase 2: zi = 1/zt. Valid iff y(R) is even and x(R) == r.
template <class Trace = NoTrace>
KGV_HD uint8_t schnorr_phase2(const fe& X, const fe& Y, const fe& zi, const fe& rx, Trace trace = Trace()) {
  fe zi2, ax, ay;
  fe_sqr(zi2, zi);
  fe_mul(ax, X, zi2);
  fe_mul(ay, Y, zi2);
  fe_mul(ay, ay, zi);
  fe_normalize(ay);
  fe_normalize(ax);
  trace(22, ax.v, 8); trace(23, ay.v, 8);
  if (ay.v[0] & 1u) return KGV_ST_INVALID;                // y(R) odd
  bool eq = true;
#pragma unroll
  for (int i = 0; i < 8; i++) eq = eq && (ax.v[i] == rx.v[i]);
  return eq ? KGV_ST_VALID : KGV_ST_INVALID;
}
// single-signature form (audit kernel, host unit tests)
template <class Tab, class GLoad, class Trace = NoTrace>
KGV_HD uint8_t schnorr_verify_core(const uint32_t* pkw, const uint32_t* mw, const uint32_t* sigw, Tab& tab, const uint32_t* gtab,
                                   GLoad gload, Trace trace = Trace()) {
  fe X, Y, zt, rx, zi;
  uint8_t st = schnorr_phase1(X, Y, zt, rx, pkw, mw, sigw, tab, gtab, gload, trace);
  if (st != KGV_ST_PENDING) return st;
  fe_inv(zi, zt);
  return schnorr_phase2(X, Y, zi, rx, trace);
}

// ECDSA verification with libsecp256k1 semantics, phase 1: parsing and range checks.
// pkw: 8 big-endian words of x, tag = first key byte.  Returns a final verdict or KGV_ST_PENDING with
// the key (qx,qy), r, s (to be inverted, possibly batched) and the reduced message m.
KGV_HD uint8_t ecdsa_phase1(fe& qx, fe& qy, uint32_t* r, uint32_t* s, uint32_t* m, uint32_t tag, const uint32_t* pkw, const uint32_t* mw,
                            const uint32_t* sigw) {
  if (tag != 2u && tag != 3u) return KGV_ST_PK_PARSE;
  limbs_from_be_words(qx.v, pkw);
  if (!fe_words_lt_p(qx.v)) return KGV_ST_PK_PARSE;
  if (!ge_lift_x(qy, qx, tag == 3u)) return KGV_ST_PK_PARSE;
  limbs_from_be_words(r, sigw);
  limbs_from_be_words(s, sigw + 8);
  if (sc_ge_n(r) || sc_ge_n(s)) return KGV_ST_SIG_PARSE;  // from_compact rejects overflow
  limbs_from_be_words(m, mw);
  sc_reduce_once(m);
  if (sc_is_high(s)) return KGV_ST_INVALID;               // verify requires low S
  if (is_zero8(r) || is_zero8(s)) return KGV_ST_INVALID;
  return KGV_ST_PENDING;
}
// phase 2: sn = s^-1 mod n.  R = (m/s)*G + (r/s)*Q, valid iff x(R) mod n == r.
template <class Tab, class GLoad>
KGV_HD uint8_t ecdsa_phase2(const fe& qx, const fe& qy, const uint32_t* r, const uint32_t* sn, const uint32_t* m, Tab& tab, const uint32_t* gtab,
                            GLoad gload) {
  uint32_t u1[8], u2[8];
  sc_mul(u1, sn, m);
  sc_mul(u2, sn, r);
  gej R;
  fe zs;
  ecmult_double(R, zs, qx, qy, u2, u1, tab, gtab, gload);
  if (R.inf) return KGV_ST_INVALID;
  // x(R) mod n == r  <=>  X == r*Zt^2  or  (r + n < p and X == (r+n)*Zt^2)
  fe zt, zt2, t, rf;
  fe_mul(zt, R.z, zs);
  fe_sqr(zt2, zt);
#pragma unroll
  for (int i = 0; i < 8; i++) rf.v[i] = r[i];
  fe_mul(t, rf, zt2);
  if (fe_equal(t, R.x)) return KGV_ST_VALID;
  const uint32_t n[8] = KGV_N_LIMBS;
  uint32_t rn[8];
  uint32_t c = add8(rn, r, n);
  if (c || !fe_words_lt_p(rn)) return KGV_ST_INVALID;     // r + n >= p
#pragma unroll
  for (int i = 0; i < 8; i++) rf.v[i] = rn[i];
  fe_mul(t, rf, zt2);
  return fe_equal(t, R.x) ? KGV_ST_VALID : KGV_ST_INVALID;
}
template <class Tab, class GLoad>
KGV_HD uint8_t ecdsa_verify_core(uint32_t tag, const uint32_t* pkw, const uint32_t* mw, const uint32_t* sigw, Tab& tab,
                                 const uint32_t* gtab, GLoad gload) {
  fe qx, qy;
  uint32_t r[8], s[8], m[8], sn[8];
  uint8_t st = ecdsa_phase1(qx, qy, r, s, m, tag, pkw, mw, sigw);
  if (st != KGV_ST_PENDING) return st;
  sc_inv(sn, s);
  return ecdsa_phase2(qx, qy, r, sn, m, tab, gtab, gload);
}

// One entry of the generator tables: v * B for v in [1, 65535], B affine; result affine.
KGV_HD void gtab_entry(fe& ox, fe& oy, uint32_t v, const fe& bx, const fe& by) {
  gej r;
  r.inf = true;
  fe_set_zero(r.x); fe_set_zero(r.y); fe_set_zero(r.z);
  for (int bit = 15; bit >= 0; bit--) {
    gej_double(r);
    if ((v >> bit) & 1u) gej_add_ge(r, bx, by);
  }
  fe zi, zi2;
  fe_inv(zi, r.z);
  fe_sqr(zi2, zi);
  fe_mul(ox, r.x, zi2);
  fe_mul(oy, r.y, zi2);
  fe_mul(oy, oy, zi);
  fe_normalize(ox);
  fe_normalize(oy);
}

#define KGV_GX_LIMBS {0x16F81798u, 0x59F2815Bu, 0x2DCE28D9u, 0x029BFCDBu, 0xCE870B07u, 0x55A06295u, 0xF9DCBBACu, 0x79BE667Eu}
#define KGV_GY_LIMBS {0xFB10D4B8u, 0x9C47D08Fu, 0xA6855419u, 0xFD17B448u, 0x0E1108A8u, 0x5DA4FBFCu, 0x26A3C465u, 0x483ADA77u}
// 2^128 * G (tools/derive_constants.py)
#define KGV_G128X_LIMBS {0x9EC4C0DAu, 0x1B7B444Cu, 0x723EA335u, 0xE88C5678u, 0x981F162Eu, 0x9239C1ADu, 0xF63B5F33u, 0x8F68B9D2u}
#define KGV_G128Y_LIMBS {0x501FFF82u, 0xF23CBF79u, 0x95510BFDu, 0xBBEA2CFEu, 0xB6BE215Du, 0xDE1D90C2u, 0xBA063986u, 0x662A9F2Du}

}  // namespace kgv
