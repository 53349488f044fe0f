// kgv_simgen.cpp — fast, seeded workload generator for the DAG-replay benchmarks (BASELINE configs 1, 3, 4, 5).
//
// Same shapes as rusty_kaspa_b200/simgen.py (which restates what simpa's miner emits, simpa/src/simulator/miner.rs:138-207:
// version-0 native transactions spending earlier outputs of the generator's own keys, P2PK Schnorr `20 <xonly> ac`, signature
// scripts `41 <sig64> 01`, committed storage mass, <= --tpb transactions per block; extended as BASELINE.json asks with 2-input
// transactions, P2PK-ECDSA and P2SH 2-of-3 multisig, and a fraction of deliberately invalid transactions), but ~100x faster than the
// Python generator so that config 3 (10 k blocks, ~1 M transactions) and config 4 (500 k transactions) can be produced inside
// bench.py.  The output is the flat batch layout of include/kgv.h, written directly.
//
// This is workload tooling (the counterpart of simpa itself), not part of the product path and not the oracle: it uses no
// elliptic-curve arithmetic at all.  Keys and nonces come as precomputed (scalar, x-coordinate) pools from
// rusty_kaspa_b200/workload.ScalarPointPool; a signature is scalar arithmetic modulo n over those pools.  Hashing (tx id, sighash,
// BIP-340 challenge) is the library's own device code compiled for the host - the same way tests/hostsim compiles it.
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../rusty_kaspa_b200/csrc/kgv_txhash.cuh"

using namespace kgv;
typedef unsigned __int128 u128;

namespace {

struct Rng {  // xoshiro256**
  uint64_t s[4];
  static uint64_t splitmix(uint64_t& x) { uint64_t z = (x += 0x9E3779B97F4A7C15ull); z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31); }
  explicit Rng(uint64_t seed) { for (auto& v : s) v = splitmix(seed); }
  static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
  uint64_t next() {
    uint64_t r = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
    s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
    return r;
  }
  uint64_t below(uint64_t n) { return (uint64_t)(((u128)next() * n) >> 64); }
  double unit() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
};

// ---- 256-bit scalars modulo the group order n, 4 x u64 little-endian
struct Sc { uint64_t v[4]; };
const Sc N_ = {{0xBFD25E8CD0364141ull, 0xBAAEDCE6AF48A03Bull, 0xFFFFFFFFFFFFFFFEull, 0xFFFFFFFFFFFFFFFFull}};
const uint64_t NC[3] = {0x402DA1732FC9BEBFull, 0x4551231950B75FC4ull, 1ull};  // 2^256 - n
bool sc_ge(const Sc& a, const Sc& b) { for (int i = 3; i >= 0; i--) { if (a.v[i] != b.v[i]) return a.v[i] > b.v[i]; } return true; }
Sc sc_sub_raw(const Sc& a, const Sc& b) { Sc r; u128 br = 0; for (int i = 0; i < 4; i++) { u128 t = (u128)a.v[i] - b.v[i] - br; r.v[i] = (uint64_t)t; br = (t >> 64) & 1; } return r; }
Sc sc_from_be(const uint8_t* p) { Sc r; for (int i = 0; i < 4; i++) { uint64_t w = 0; for (int k = 0; k < 8; k++) w = (w << 8) | p[8 * (3 - i) + k]; r.v[i] = w; } return r; }
void sc_to_be(uint8_t* p, const Sc& a) { for (int i = 0; i < 4; i++) for (int k = 0; k < 8; k++) p[8 * (3 - i) + k] = (uint8_t)(a.v[i] >> (56 - 8 * k)); }
Sc sc_reduce_once(Sc a) { return sc_ge(a, N_) ? sc_sub_raw(a, N_) : a; }
Sc sc_add(const Sc& a, const Sc& b) {
  Sc r; u128 c = 0;
  for (int i = 0; i < 4; i++) { c += (u128)a.v[i] + b.v[i]; r.v[i] = (uint64_t)c; c >>= 64; }
  if (c) { u128 d = 0; for (int i = 0; i < 4; i++) { d += (u128)r.v[i] + (i < 3 ? NC[i] : 0); r.v[i] = (uint64_t)d; d >>= 64; } }  // + (2^256 mod n)
  return sc_reduce_once(r);
}
// t[0..len) += x[0..nx) * NC, little-endian u64 limbs (no overflow past len by construction)
void fold(uint64_t* acc, int len, const uint64_t* x, int nx) {
  for (int i = 0; i < nx; i++) {
    u128 c = 0;
    for (int j = 0; j < 3 && i + j < len; j++) { c += (u128)x[i] * NC[j] + acc[i + j]; acc[i + j] = (uint64_t)c; c >>= 64; }
    for (int k = i + 3; k < len && c; k++) { c += acc[k]; acc[k] = (uint64_t)c; c >>= 64; }
  }
}
Sc sc_mul(const Sc& a, const Sc& b) {
  uint64_t t[8] = {0};
  for (int i = 0; i < 4; i++) { u128 c = 0; for (int j = 0; j < 4; j++) { c += (u128)a.v[i] * b.v[j] + t[i + j]; t[i + j] = (uint64_t)c; c >>= 64; } t[i + 4] = (uint64_t)c; }
  uint64_t u[7] = {t[0], t[1], t[2], t[3], 0, 0, 0};
  fold(u, 7, t + 4, 4);            // < 2^386
  uint64_t w[5] = {u[0], u[1], u[2], u[3], 0};
  fold(w, 5, u + 4, 3);            // < 2^260
  uint64_t z[5] = {w[0], w[1], w[2], w[3], 0};
  fold(z, 5, w + 4, 1);            // < 2^256 + 2^134
  Sc r = {{z[0], z[1], z[2], z[3]}};
  if (z[4]) { u128 d = 0; for (int i = 0; i < 4; i++) { d += (u128)r.v[i] + (i < 3 ? NC[i] : 0); r.v[i] = (uint64_t)d; d >>= 64; } }
  return sc_reduce_once(sc_reduce_once(r));
}
bool sc_is_high(const Sc& a) {
  const Sc half = {{0xDFE92F46681B20A0ull, 0x5D576E7357A4501Dull, 0xFFFFFFFFFFFFFFFFull, 0x7FFFFFFFFFFFFFFFull}};
  return !sc_ge(half, a);
}

enum { KIND_P2PK = 0, KIND_P2PK_ECDSA = 1, KIND_MS = 2, KIND_MS_ECDSA = 3 };

struct Config {
  uint64_t seed;
  uint32_t n_keys, n_nonces;
  uint64_t storage_mass_parameter, coinbase_maturity;
  double mix[4];
  double frac_two_inputs, frac_invalid;
  uint32_t coinbase_outputs;
  uint32_t pad_;
  uint64_t subsidy;
};

struct Utxo {
  uint8_t txid[32];
  uint32_t index;
  uint64_t amount;
  uint32_t script_off, script_len;  // in bytes
  uint8_t kind;
  uint32_t keys[3];
  uint32_t redeem_off, redeem_len;
  uint64_t daa;
  bool coinbase;
  uint64_t touched_block;
};

struct Gen {
  Config cfg;
  Rng rng;
  std::vector<Sc> key_d, nonce_k, nonce_kinv;
  std::vector<uint8_t> key_x, nonce_x;  // 32 bytes each, big-endian
  std::vector<kgv_tx> txs;
  std::vector<kgv_input> inputs;
  std::vector<kgv_output> outputs;
  std::vector<uint8_t> bytes;
  std::vector<uint32_t> block_first_tx;  // n_blocks + 1
  std::vector<uint64_t> block_pov;
  std::vector<Utxo> utxos;
  uint64_t daa = 0, n_signatures = 0, n_invalid = 0;
  double mix_cdf[4];
  explicit Gen(const Config& c) : cfg(c), rng(c.seed) {
    double s = 0;
    for (int i = 0; i < 4; i++) s += c.mix[i];
    double a = 0;
    for (int i = 0; i < 4; i++) { a += c.mix[i] / s; mix_cdf[i] = a; }
    block_first_tx.push_back(0);
  }
  uint32_t put(const uint8_t* p, size_t n) { uint32_t off = (uint32_t)bytes.size(); bytes.insert(bytes.end(), p, p + n); return off; }

  struct NewScript { uint8_t kind; uint32_t keys[3]; uint32_t redeem_off, redeem_len; uint8_t spk[35]; uint32_t spk_len; };
  NewScript new_output_script() {
    NewScript o{};
    double u = rng.unit();
    int kind = 0;
    while (kind < 3 && u >= mix_cdf[kind]) kind++;
    o.kind = (uint8_t)kind;
    if (kind == KIND_P2PK) {
      uint32_t k = (uint32_t)rng.below(cfg.n_keys);
      o.keys[0] = k; o.spk[0] = 0x20; memcpy(o.spk + 1, &key_x[32 * k], 32); o.spk[33] = 0xAC; o.spk_len = 34;
      return o;
    }
    if (kind == KIND_P2PK_ECDSA) {
      uint32_t k = (uint32_t)rng.below(cfg.n_keys);
      o.keys[0] = k; o.spk[0] = 0x21; o.spk[1] = 0x02; memcpy(o.spk + 2, &key_x[32 * k], 32); o.spk[34] = 0xAB; o.spk_len = 35;
      return o;
    }
    // three distinct keys
    uint32_t a = (uint32_t)rng.below(cfg.n_keys), b, c;
    do b = (uint32_t)rng.below(cfg.n_keys); while (b == a);
    do c = (uint32_t)rng.below(cfg.n_keys); while (c == a || c == b);
    o.keys[0] = a; o.keys[1] = b; o.keys[2] = c;
    uint8_t redeem[110];
    uint32_t n = 0;
    redeem[n++] = 0x52;
    for (int j = 0; j < 3; j++) {
      if (kind == KIND_MS) { redeem[n++] = 0x20; }
      else { redeem[n++] = 0x21; redeem[n++] = 0x02; }
      memcpy(redeem + n, &key_x[32 * o.keys[j]], 32); n += 32;
    }
    redeem[n++] = 0x53;
    redeem[n++] = kind == KIND_MS ? 0xAE : 0xA9;
    o.redeem_off = put(redeem, n); o.redeem_len = n;
    Blake2b h; b2b_init(h, B2B_UNKEYED); b2b_bytes(h, redeem, n);
    uint64_t d[4]; b2b_final(h, d);
    o.spk[0] = 0xAA; o.spk[1] = 0x20; memcpy(o.spk + 2, d, 32); o.spk[34] = 0x87; o.spk_len = 35;
    return o;
  }

  static void be_words(uint32_t* w, const uint8_t* p) { for (int i = 0; i < 8; i++) w[i] = ((uint32_t)p[4 * i] << 24) | ((uint32_t)p[4 * i + 1] << 16) | ((uint32_t)p[4 * i + 2] << 8) | p[4 * i + 3]; }
  // msg: 8 big-endian words. sig out: 64 bytes
  void sign(uint8_t* sig, uint32_t key_idx, const uint32_t* mw, bool ecdsa) {
    uint32_t j = (uint32_t)rng.below(cfg.n_nonces);
    n_signatures++;
    const uint8_t* rx = &nonce_x[32 * j];
    uint8_t mbytes[32];
    for (int i = 0; i < 8; i++) { mbytes[4 * i] = mw[i] >> 24; mbytes[4 * i + 1] = mw[i] >> 16; mbytes[4 * i + 2] = mw[i] >> 8; mbytes[4 * i + 3] = mw[i]; }
    if (!ecdsa) {
      uint32_t rw[8], pw[8], ew[8];
      be_words(rw, rx); be_words(pw, &key_x[32 * key_idx]);
      bip340_challenge(ew, rw, pw, mw);
      uint8_t eb[32];
      for (int i = 0; i < 8; i++) { eb[4 * i] = ew[i] >> 24; eb[4 * i + 1] = ew[i] >> 16; eb[4 * i + 2] = ew[i] >> 8; eb[4 * i + 3] = ew[i]; }
      Sc e = sc_reduce_once(sc_from_be(eb));
      Sc s = sc_add(nonce_k[j], sc_mul(e, key_d[key_idx]));
      memcpy(sig, rx, 32);
      sc_to_be(sig + 32, s);
      return;
    }
    Sc r = sc_reduce_once(sc_from_be(rx));
    Sc m = sc_reduce_once(sc_from_be(mbytes));
    Sc s = sc_mul(nonce_kinv[j], sc_add(m, sc_mul(r, key_d[key_idx])));
    if (sc_is_high(s)) s = sc_sub_raw(N_, s);
    sc_to_be(sig, r);
    sc_to_be(sig + 32, s);
  }

  static uint64_t plurality(uint32_t script_len) { return (63 + (uint64_t)script_len + 99) / 100; }
  // consensus/core/src/mass/mod.rs:338-410 for generator-sized values (no overflow possible)
  uint64_t storage_mass(const uint64_t* in_amt, const uint32_t* in_len, int n_in, const uint64_t* out_val, const uint32_t* out_len, int n_out) const {
    const uint64_t C = cfg.storage_mass_parameter;
    uint64_t outs_plur = 0, harm_outs = 0, ins_plur = 0;
    for (int i = 0; i < n_out; i++) { uint64_t p = plurality(out_len[i]); outs_plur += p; harm_outs += C * p * p / out_val[i]; }
    for (int i = 0; i < n_in; i++) ins_plur += plurality(in_len[i]);
    if (outs_plur == 1 || (n_in <= 2 && (ins_plur == 1 || (outs_plur == 2 && ins_plur == 2)))) {
      uint64_t harm_ins = 0;
      for (int i = 0; i < n_in; i++) { uint64_t p = plurality(in_len[i]); harm_ins += C * p * p / in_amt[i]; }
      return harm_outs > harm_ins ? harm_outs - harm_ins : 0;
    }
    uint64_t sum = 0;
    for (int i = 0; i < n_in; i++) sum += in_amt[i];
    uint64_t mean = sum / ins_plur;
    uint64_t arith = ins_plur * (C / mean);
    return harm_outs > arith ? harm_outs - arith : 0;
  }

  bool pick(Utxo& out, uint64_t pov) {
    for (int tries = 0; tries < 50; tries++) {
      if (utxos.empty()) return false;
      size_t i = (size_t)rng.below(utxos.size());
      const Utxo& u = utxos[i];
      if (u.coinbase && u.daa + cfg.coinbase_maturity > pov) continue;
      if (u.amount < 4 || u.touched_block == pov) continue;
      out = u;
      utxos[i] = utxos.back();
      utxos.pop_back();
      out.touched_block = pov;
      return true;
    }
    return false;
  }

  void add_output_record(uint64_t value, const NewScript& ns) {
    kgv_output o{};
    o.value = value; o.script_off = put(ns.spk, ns.spk_len); o.script_len = ns.spk_len; o.spk_version = 0;
    outputs.push_back(o);
  }
  Utxo utxo_of(const uint8_t* txid, uint32_t index, const kgv_output& o, const NewScript& ns, uint64_t daa_, bool cb) {
    Utxo u{};
    memcpy(u.txid, txid, 32); u.index = index; u.amount = o.value; u.script_off = o.script_off; u.script_len = o.script_len; u.kind = ns.kind;
    memcpy(u.keys, ns.keys, sizeof u.keys); u.redeem_off = ns.redeem_off; u.redeem_len = ns.redeem_len; u.daa = daa_; u.coinbase = cb; u.touched_block = 0;
    return u;
  }
  BatchView view(const DevEntry* ent_base) const { return BatchView{txs.data(), inputs.data(), outputs.data(), ent_base, bytes.data()}; }

  void make_block(uint32_t n_txs) {
    daa++;
    const uint64_t pov = daa;
    std::vector<Utxo> fresh;
    // coinbase
    {
      kgv_tx t{};
      t.first_input = (uint32_t)inputs.size(); t.n_inputs = 0; t.first_output = (uint32_t)outputs.size(); t.n_outputs = cfg.coinbase_outputs;
      t.version = 0; t.subnetwork_id[0] = 1; t.flags = 1;
      uint8_t payload[15];
      for (int i = 0; i < 8; i++) payload[i] = (uint8_t)(pov >> (8 * i));
      memcpy(payload + 8, "kgv-sim", 7);
      t.payload_off = put(payload, 15); t.payload_len = 15;
      std::vector<NewScript> created;
      for (uint32_t i = 0; i < cfg.coinbase_outputs; i++) { NewScript ns = new_output_script(); add_output_record(cfg.subsidy / cfg.coinbase_outputs, ns); created.push_back(ns); }
      txs.push_back(t);
      uint64_t id[4];
      tx_id(id, view(nullptr), (uint32_t)txs.size() - 1);
      for (uint32_t i = 0; i < cfg.coinbase_outputs; i++) fresh.push_back(utxo_of((const uint8_t*)id, i, outputs[t.first_output + i], created[i], pov, true));
    }
    for (uint32_t k = 0; k < n_txs; k++) make_tx(pov, fresh);
    utxos.insert(utxos.end(), fresh.begin(), fresh.end());
    block_first_tx.push_back((uint32_t)txs.size());
    block_pov.push_back(pov);
  }

  void make_tx(uint64_t pov, std::vector<Utxo>& fresh) {
    int want = rng.unit() < cfg.frac_two_inputs ? 2 : 1;
    Utxo ins[2];
    int n_in = 0;
    for (int i = 0; i < want; i++) if (pick(ins[n_in], pov)) n_in++;
    if (n_in == 0) return;
    uint64_t total = 0;
    for (int i = 0; i < n_in; i++) total += ins[i].amount;
    const uint64_t fee = 1;
    kgv_tx t{};
    t.first_input = (uint32_t)inputs.size(); t.n_inputs = (uint32_t)n_in; t.first_output = (uint32_t)outputs.size(); t.n_outputs = 2;
    t.payload_off = (uint32_t)bytes.size(); t.payload_len = 0;
    NewScript created[2];
    uint64_t v0 = (total - fee) / 2, vals[2] = {v0, total - fee - v0};
    for (int i = 0; i < 2; i++) { created[i] = new_output_script(); add_output_record(vals[i], created[i]); }
    for (int i = 0; i < n_in; i++) {
      kgv_input in{};
      memcpy(in.prev_txid, ins[i].txid, 32); in.prev_index = ins[i].index; in.sequence = 0;
      in.sig_op_count = (ins[i].kind == KIND_P2PK || ins[i].kind == KIND_P2PK_ECDSA) ? 1 : 3;
      inputs.push_back(in);
    }
    uint64_t in_amt[2]; uint32_t in_len[2], out_len[2];
    for (int i = 0; i < n_in; i++) { in_amt[i] = ins[i].amount; in_len[i] = ins[i].script_len; }
    for (int i = 0; i < 2; i++) out_len[i] = created[i].spk_len;
    t.mass = storage_mass(in_amt, in_len, n_in, vals, out_len, 2);
    const bool invalid = rng.unit() < cfg.frac_invalid;
    int mode = invalid ? (int)rng.below(8) : -1;
    const int k0 = ins[0].kind;
    if ((mode == 5 && k0 != KIND_P2PK_ECDSA) || (mode == 7 && k0 != KIND_MS && k0 != KIND_MS_ECDSA) || (mode == 6 && (k0 == KIND_MS || k0 == KIND_MS_ECDSA))) mode = 4;
    if (mode == 0) t.mass += 1;                                               // WrongMass
    else if (mode == 1) outputs[t.first_output].value += total;               // SpendTooHigh
    else if (mode == 2) for (int i = 0; i < 32; i++) inputs[t.first_input].prev_txid[i] = (uint8_t)rng.next();  // MissingTxOutpoints
    else if (mode == 3) inputs[t.first_input].sig_op_count = 0;               // ExceededSigOpLimit
    txs.push_back(t);
    const uint32_t ti = (uint32_t)txs.size() - 1;
    // signatures
    DevEntry ent[2];
    for (int i = 0; i < n_in; i++) ent[i] = DevEntry{ins[i].amount, ins[i].daa, nullptr, ins[i].script_len, 0, (uint8_t)ins[i].coinbase, 1};
    std::vector<uint8_t> sigscripts[2];
    {
      // sighash reads spk bytes through ent[].script: the arena may not grow while the pointers are live
      for (int i = 0; i < n_in; i++) ent[i].script = bytes.data() + ins[i].script_off;
      BatchView v = view(ent - t.first_input);
      SigHashReused reu;
      sighash_reused(reu, v, ti);
      for (int i = 0; i < n_in; i++) {
        const Utxo& u = ins[i];
        const bool ecdsa = u.kind == KIND_P2PK_ECDSA || u.kind == KIND_MS_ECDSA;
        uint32_t mw[8];
        sighash_final(mw, v, ti, t.first_input + i, 1, ecdsa, reu);
        std::vector<uint8_t>& ss = sigscripts[i];
        if (u.kind == KIND_P2PK || u.kind == KIND_P2PK_ECDSA) {
          uint8_t sig[64];
          sign(sig, u.keys[0], mw, ecdsa);
          if (mode == 4 && i == 0) sig[40] ^= 1;                               // EvalFalse
          if (mode == 5 && i == 0 && ecdsa) { Sc s = sc_from_be(sig + 32); s = sc_sub_raw(N_, s); sc_to_be(sig + 32, s); }  // high S
          ss.push_back(0x41); ss.insert(ss.end(), sig, sig + 64); ss.push_back((mode == 6 && i == 0) ? 0x03 : 0x01);
        } else {
          int drop = (int)rng.below(3);  // the signer pair = {0,1,2} minus one, in key order
          int pair[2], n = 0;
          for (int q = 0; q < 3; q++) if (q != drop) pair[n++] = q;
          if (mode == 7 && i == 0) { int x = pair[0]; pair[0] = pair[1]; pair[1] = x; }  // wrong order => NullFail
          for (int q = 0; q < 2; q++) {
            uint8_t sig[64];
            sign(sig, u.keys[pair[q]], mw, ecdsa);
            if (mode == 4 && i == 0 && q == 1) sig[40] ^= 1;
            ss.push_back(0x41); ss.insert(ss.end(), sig, sig + 64); ss.push_back(0x01);
          }
          // push(redeem): 102 / 105 bytes => OP_PUSHDATA1
          ss.push_back(0x4C); ss.push_back((uint8_t)u.redeem_len);
          ss.insert(ss.end(), bytes.begin() + u.redeem_off, bytes.begin() + u.redeem_off + u.redeem_len);
        }
      }
    }
    for (int i = 0; i < n_in; i++) {
      inputs[t.first_input + i].sigscript_off = put(sigscripts[i].data(), sigscripts[i].size());
      inputs[t.first_input + i].sigscript_len = (uint32_t)sigscripts[i].size();
    }
    if (mode == -1) {
      uint64_t id[4];
      tx_id(id, view(nullptr), ti);
      for (uint32_t i = 0; i < 2; i++) fresh.push_back(utxo_of((const uint8_t*)id, i, outputs[t.first_output + i], created[i], pov, false));
    } else {
      n_invalid++;
      for (int i = 0; i < n_in; i++) utxos.push_back(ins[i]);  // not accepted: its inputs stay unspent (and untouchable for the rest of this block)
    }
  }
};

}  // namespace

extern "C" {

void* sg_create(const Config* cfg, const uint8_t* key_scalars, const uint8_t* key_xs, const uint8_t* nonce_scalars, const uint8_t* nonce_kinv, const uint8_t* nonce_xs) {
  Gen* g = new Gen(*cfg);
  for (uint32_t i = 0; i < cfg->n_keys; i++) g->key_d.push_back(sc_from_be(key_scalars + 32 * i));
  g->key_x.assign(key_xs, key_xs + 32 * (size_t)cfg->n_keys);
  for (uint32_t i = 0; i < cfg->n_nonces; i++) { g->nonce_k.push_back(sc_from_be(nonce_scalars + 32 * i)); g->nonce_kinv.push_back(sc_from_be(nonce_kinv + 32 * i)); }
  g->nonce_x.assign(nonce_xs, nonce_xs + 32 * (size_t)cfg->n_nonces);
  return g;
}
void sg_destroy(void* h) { delete (Gen*)h; }
void sg_generate(void* h, uint32_t n_blocks, uint32_t txs_per_block) {
  Gen* g = (Gen*)h;
  for (uint32_t b = 0; b < n_blocks; b++) g->make_block(txs_per_block);
}
// counts: [n_txs, n_inputs, n_outputs, n_bytes, n_blocks, n_signatures, n_invalid, n_utxos]
void sg_counts(void* h, uint64_t* out) {
  Gen* g = (Gen*)h;
  out[0] = g->txs.size(); out[1] = g->inputs.size(); out[2] = g->outputs.size(); out[3] = g->bytes.size(); out[4] = g->block_pov.size();
  out[5] = g->n_signatures; out[6] = g->n_invalid; out[7] = g->utxos.size();
}
// copies everything generated so far into caller arrays and leaves the generator ready to continue (the arrays keep growing:
// a later call returns the whole history again; use sg_reset_output between windows to start a fresh batch)
void sg_copy(void* h, kgv_tx* txs, kgv_input* inputs, kgv_output* outputs, uint8_t* bytes, uint32_t* block_first_tx, uint64_t* block_pov) {
  Gen* g = (Gen*)h;
  memcpy(txs, g->txs.data(), g->txs.size() * sizeof(kgv_tx));
  memcpy(inputs, g->inputs.data(), g->inputs.size() * sizeof(kgv_input));
  memcpy(outputs, g->outputs.data(), g->outputs.size() * sizeof(kgv_output));
  memcpy(bytes, g->bytes.data(), g->bytes.size());
  memcpy(block_first_tx, g->block_first_tx.data(), g->block_first_tx.size() * 4);
  memcpy(block_pov, g->block_pov.data(), g->block_pov.size() * 8);
}
// drops the emitted batch (keeps the spendable outputs, whose scripts are copied into the new arena) so that the next blocks form
// a new, self-contained batch: windows of a long replay are generated one after the other with bounded memory
void sg_reset_output(void* h) {
  Gen* g = (Gen*)h;
  std::vector<uint8_t> nb;
  for (Utxo& u : g->utxos) {
    uint32_t so = (uint32_t)nb.size();
    nb.insert(nb.end(), g->bytes.begin() + u.script_off, g->bytes.begin() + u.script_off + u.script_len);
    u.script_off = so;
    if (u.redeem_len) {
      uint32_t ro = (uint32_t)nb.size();
      nb.insert(nb.end(), g->bytes.begin() + u.redeem_off, g->bytes.begin() + u.redeem_off + u.redeem_len);
      u.redeem_off = ro;
    }
  }
  g->bytes.swap(nb);
  g->txs.clear(); g->inputs.clear(); g->outputs.clear();
  g->block_first_tx.assign(1, 0u);
  g->block_pov.clear();
}

}  // extern "C"
