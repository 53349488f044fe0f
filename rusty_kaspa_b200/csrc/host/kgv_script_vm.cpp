// kgv_script_vm.cpp — see kgv_script_vm.h.  Plain C++ (host code of libkgv.so).
#include "kgv_script_vm.h"

#include <cstring>

// OpSHA256 / OpBlake2b hash stack items on the host; the hashing headers are written to compile for both
// sides (device bodies under nvcc, plain C++ here).
#include "../kgv_blake2b.cuh"
#include "../kgv_sha256.cuh"

namespace kgv_host {

using Bytes = std::vector<uint8_t>;
using Stack = std::vector<Bytes>;

static const size_t MAX_STACK_SIZE = 244, MAX_SCRIPTS_SIZE = 10000, MAX_SCRIPT_ELEMENT_SIZE = 520;  // lib.rs:41-54
static const int MAX_OPS_PER_SCRIPT = 201, MAX_PUB_KEYS_PER_MULTISIG = 20;
static const uint64_t LOCK_TIME_THRESHOLD = 500000000000ull, SEQUENCE_LOCK_TIME_DISABLED = 1ull << 63, SEQUENCE_LOCK_TIME_MASK = 0xffffffffull;
static const uint8_t NO_COST_OPCODE = 0x60;

// ---------------------------------------------------------------------------------------------- data_stack.rs
static bool as_bool(const Bytes& v) {  // :206-214
  if (v.empty()) return false;
  if ((v.back() & 0x7f) != 0) return true;
  for (size_t i = 0; i + 1 < v.size(); i++)
    if (v[i]) return true;
  return false;
}
static Bytes serialize_i64(int64_t x) {  // :109-137
  Bytes out;
  bool neg = x < 0;
  uint64_t p = neg ? (uint64_t)0 - (uint64_t)x : (uint64_t)x;
  bool last_sat = false;
  while (p) {
    uint8_t b = (uint8_t)(p & 0xff);
    last_sat = (b & 0x80) != 0;
    out.push_back(b);
    p >>= 8;
  }
  if (last_sat) out.push_back(0);
  if (neg) out.back() |= 0x80;
  return out;
}
// SizedEncodeInt<LEN>::deserialize (:177-190) -> error or value
static ScriptErr deserialize_num(const Bytes& v, size_t maxlen, int64_t& out) {
  if (v.size() > maxlen) return SERR_NUMBER_TOO_BIG;
  if (v.size() > 8) return SERR_NOT_MINIMAL_DATA;
  if (v.empty()) { out = 0; return SERR_OK; }
  if ((v.back() & 0x7f) == 0) {  // check_minimal_data_encoding :87-107
    if (v.size() == 1 || (v[v.size() - 2] & 0x80) == 0) return SERR_NOT_MINIMAL_DATA;
  }
  uint8_t msb = v.back();
  int64_t acc = msb & 0x7f;
  for (size_t i = v.size() - 1; i-- > 0;) acc = (int64_t)(((uint64_t)acc << 8) + v[i]);
  out = (msb & 0x80) ? -acc : acc;
  return SERR_OK;
}

struct Engine {
  const kgv_tx_batch& b;
  uint32_t tx, idx;           // idx = input index within the tx
  const kgv_tx& t;
  const kgv_input& in;
  const kgv_utxo_entry& entry;
  const VerdictFn& verdict;
  std::vector<SigRequest>* missing;
  Stack dstack, astack;
  std::vector<int> cond;      // 0 False, 1 True, 2 Skip
  int num_ops = 0;
  int sigops_remaining;

  Engine(const kgv_tx_batch& b_, uint32_t tx_, uint32_t idx_, const VerdictFn& v, std::vector<SigRequest>* m)
      : b(b_), tx(tx_), idx(idx_), t(b_.txs[tx_]), in(b_.inputs[b_.txs[tx_].first_input + idx_]), entry(b_.entries[b_.txs[tx_].first_input + idx_]),
        verdict(v), missing(m), sigops_remaining(in.sig_op_count) {}

  bool executing() const { return cond.empty() || cond.back() == 1; }

  // ---- stack helpers (data_stack.rs:216-330)
  ScriptErr pop_raw(Bytes& a) {
    if (dstack.size() < 1) return SERR_INVALID_STACK_OPERATION;
    a = std::move(dstack.back()); dstack.pop_back();
    return SERR_OK;
  }
  ScriptErr pop_nums(int n, size_t maxlen, int64_t* out) {
    if ((int)dstack.size() < n) return SERR_INVALID_STACK_OPERATION;
    Stack items(dstack.end() - n, dstack.end());
    dstack.resize(dstack.size() - n);
    for (int i = 0; i < n; i++) { ScriptErr e = deserialize_num(items[i], maxlen, out[i]); if (e) return e; }
    return SERR_OK;
  }
  ScriptErr pop_bool(bool& v) {
    if (dstack.empty()) return SERR_INVALID_STACK_OPERATION;
    v = as_bool(dstack.back()); dstack.pop_back();
    return SERR_OK;
  }
  ScriptErr push_num(int64_t x) {
    Bytes s = serialize_i64(x);
    if (s.size() > 8) return SERR_SERIALIZATION;  // SerializationError::NumberTooLong
    dstack.push_back(std::move(s));
    return SERR_OK;
  }
  void push_bool(bool v) { dstack.push_back(v ? Bytes{1} : Bytes{}); }

  // ---- signature checks (lib.rs:574-643): verdicts come from the GPU batch
  ScriptErr check_sig(uint8_t hash_type, const Bytes& key, const uint8_t* sig, size_t siglen, bool ecdsa, bool& valid) {
    if (sigops_remaining == 0) return SERR_EXCEEDED_SIGOP_LIMIT;
    sigops_remaining--;
    if (siglen != 64) return SERR_SIG_LENGTH;
    if (key.size() != (ecdsa ? 33u : 32u)) return SERR_PUBKEY_FORMAT;
    SigRequest rq;
    rq.tx = tx; rq.input_abs = t.first_input + idx; rq.hash_type = hash_type; rq.ecdsa = ecdsa ? 1 : 0; rq.key = key;
    memcpy(rq.sig, sig, 64);
    int v = verdict(rq);
    if (v < 0) { if (missing) missing->push_back(rq); return SERR_NEEDS_SIG_VERDICTS; }
    if (v == KGV_SIG_PK_PARSE_ERR || v == KGV_SIG_SIG_PARSE_ERR) return SERR_INVALID_SIGNATURE;
    valid = v == KGV_SIG_VALID;
    return SERR_OK;
  }
  static bool sighash_type_ok(uint8_t t) { return t == 1 || t == 2 || t == 4 || t == 0x81 || t == 0x82 || t == 0x84; }

  ScriptErr op_checksig(bool ecdsa) {  // opcodes/mod.rs:746-790
    if (dstack.size() < 2) return SERR_INVALID_STACK_OPERATION;
    Bytes key = std::move(dstack.back()); dstack.pop_back();
    Bytes sig = std::move(dstack.back()); dstack.pop_back();
    if (sig.empty()) { push_bool(false); return SERR_OK; }
    uint8_t typ = sig.back(); sig.pop_back();
    if (!sighash_type_ok(typ)) return SERR_INVALID_SIGHASH_TYPE;
    bool valid = false;
    ScriptErr e = check_sig(typ, key, sig.data(), sig.size(), ecdsa, valid);
    if (e) return e;
    push_bool(valid);
    return SERR_OK;
  }

  ScriptErr op_checkmultisig(bool ecdsa) {  // lib.rs:488-571
    int64_t nk;
    ScriptErr e = pop_nums(1, 4, &nk);
    if (e) return e;
    if (nk < 0 || nk > MAX_PUB_KEYS_PER_MULTISIG) return SERR_INVALID_PUBKEY_COUNT;
    num_ops += (int)nk;
    if (num_ops > MAX_OPS_PER_SCRIPT) return SERR_TOO_MANY_OPERATIONS;
    if (dstack.size() < (size_t)nk) return SERR_INVALID_STACK_OPERATION;
    Stack keys(dstack.end() - nk, dstack.end());
    dstack.resize(dstack.size() - nk);
    int64_t ns;
    e = pop_nums(1, 4, &ns);
    if (e) return e;
    if (ns < 0 || ns > nk) return SERR_INVALID_SIGNATURE_COUNT;
    if (dstack.size() < (size_t)ns) return SERR_INVALID_STACK_OPERATION;
    Stack sigs(dstack.end() - ns, dstack.end());
    dstack.resize(dstack.size() - ns);
    bool failed = false;
    size_t ki = 0;
    for (size_t si = 0; si < sigs.size() && !failed; si++) {
      const Bytes& s = sigs[si];
      if (s.empty()) { failed = true; break; }
      uint8_t typ = s.back();
      if (!sighash_type_ok(typ)) return SERR_INVALID_SIGHASH_TYPE;
      for (;;) {
        if (keys.size() - ki < sigs.size() - si) { failed = true; break; }
        const Bytes& key = keys[ki++];
        bool valid = false;
        e = check_sig(typ, key, s.data(), s.size() - 1, ecdsa, valid);
        if (e) return e;
        if (valid) break;
      }
    }
    if (failed) {
      for (const Bytes& s : sigs)
        if (!s.empty()) return SERR_NULL_FAIL;
    }
    push_bool(!failed);
    return SERR_OK;
  }

  // ---- minimal push rule (opcodes/mod.rs:141-190)
  static ScriptErr check_minimal_push(uint8_t op, const uint8_t* d, size_t n) {
    if (n == 0) return op != 0x00 ? SERR_NOT_MINIMAL_DATA : SERR_OK;
    if (n == 1 && d[0] >= 1 && d[0] <= 16) return op != 0x51 + d[0] - 1 ? SERR_NOT_MINIMAL_DATA : SERR_OK;
    if (n == 1 && d[0] == 0x81) return op != 0x4f ? SERR_NOT_MINIMAL_DATA : SERR_OK;
    if (n <= 75) return op != n ? SERR_NOT_MINIMAL_DATA : SERR_OK;
    if (n <= 255) return op != 0x4c ? SERR_NOT_MINIMAL_DATA : SERR_OK;
    if (n < 65535 && op != 0x4d) return SERR_NOT_MINIMAL_DATA;
    return SERR_OK;
  }
  static bool is_disabled(uint8_t op) {  // :103-123
    switch (op) { case 0x7e: case 0x7f: case 0x80: case 0x81: case 0x83: case 0x84: case 0x85: case 0x86: case 0x8d: case 0x8e:
                  case 0x95: case 0x96: case 0x97: case 0x98: case 0x99: return true; default: return false; }
  }

  Bytes spk_to_bytes(uint16_t version, const uint8_t* script, size_t n) {  // lib.rs:645-653 (version big-endian || script)
    Bytes v;
    v.push_back((uint8_t)(version >> 8)); v.push_back((uint8_t)version);
    v.insert(v.end(), script, script + n);
    return v;
  }

  ScriptErr exec(uint8_t op, const uint8_t* data, size_t dlen) {
    int64_t n[3];
    ScriptErr e;
    Bytes a;
    if (op == 0x00) { dstack.emplace_back(); return SERR_OK; }
    if (op <= 0x4e) { dstack.emplace_back(data, data + dlen); return SERR_OK; }
    if (op == 0x4f) return push_num(-1);
    if (op >= 0x51 && op <= 0x60) return push_num(op - 0x50);
    switch (op) {
      case 0x50: case 0x62: case 0x65: case 0x66: case 0x89: case 0x8a: return SERR_OPCODE_RESERVED;
      case 0x61: return SERR_OK;
      case 0x63: case 0x64: {  // OpIf / OpNotIf
        int c = 2;
        if (executing()) {
          if (dstack.empty()) return SERR_EMPTY_STACK;
          Bytes buf = std::move(dstack.back()); dstack.pop_back();
          if (buf.size() > 1) return SERR_EXPECTED_BOOLEAN;
          bool truth;
          if (buf.empty()) truth = false;
          else if (buf[0] == 1) truth = true;
          else return SERR_EXPECTED_BOOLEAN;
          c = (truth == (op == 0x63)) ? 1 : 0;
        }
        cond.push_back(c);
        return SERR_OK;
      }
      case 0x67: if (cond.empty()) return SERR_COND_STACK_EMPTY; if (cond.back() != 2) cond.back() ^= 1; return SERR_OK;
      case 0x68: if (cond.empty()) return SERR_COND_STACK_EMPTY; cond.pop_back(); return SERR_OK;
      case 0x69: { bool v; if ((e = pop_bool(v))) return e; return v ? SERR_OK : SERR_VERIFY; }
      case 0x6a: return SERR_EARLY_RETURN;
      case 0x6b: if ((e = pop_raw(a))) return e; astack.push_back(std::move(a)); return SERR_OK;
      case 0x6c: if (astack.empty()) return SERR_EMPTY_STACK; dstack.push_back(std::move(astack.back())); astack.pop_back(); return SERR_OK;
      case 0x6d: if (dstack.size() < 2) return SERR_INVALID_STACK_OPERATION; dstack.resize(dstack.size() - 2); return SERR_OK;
      case 0x6e: case 0x6f: case 0x76: {  // 2DUP 3DUP DUP
        size_t k = op == 0x6e ? 2 : op == 0x6f ? 3 : 1;
        if (dstack.size() < k) return SERR_INVALID_STACK_OPERATION;
        for (size_t i = 0; i < k; i++) { Bytes c = dstack[dstack.size() - k]; dstack.push_back(std::move(c)); }
        return SERR_OK;
      }
      case 0x70: case 0x78: {  // 2OVER OVER
        size_t k = op == 0x70 ? 2 : 1;
        if (dstack.size() < 2 * k) return SERR_INVALID_STACK_OPERATION;
        for (size_t i = 0; i < k; i++) { Bytes c = dstack[dstack.size() - 2 * k]; dstack.push_back(std::move(c)); }
        return SERR_OK;
      }
      case 0x71: case 0x7b: {  // 2ROT ROT
        size_t k = op == 0x71 ? 2 : 1;
        if (dstack.size() < 3 * k) return SERR_INVALID_STACK_OPERATION;
        Stack d(dstack.end() - 3 * k, dstack.end() - 2 * k);
        dstack.erase(dstack.end() - 3 * k, dstack.end() - 2 * k);
        dstack.insert(dstack.end(), d.begin(), d.end());
        return SERR_OK;
      }
      case 0x72: case 0x7c: {  // 2SWAP SWAP
        size_t k = op == 0x72 ? 2 : 1;
        if (dstack.size() < 2 * k) return SERR_INVALID_STACK_OPERATION;
        Stack d(dstack.end() - 2 * k, dstack.end() - k);
        dstack.erase(dstack.end() - 2 * k, dstack.end() - k);
        dstack.insert(dstack.end(), d.begin(), d.end());
        return SERR_OK;
      }
      case 0x73: if (dstack.empty()) return SERR_INVALID_STACK_OPERATION; if (as_bool(dstack.back())) { Bytes c = dstack.back(); dstack.push_back(std::move(c)); } return SERR_OK;
      case 0x74: return push_num((int64_t)dstack.size());
      case 0x75: if (dstack.empty()) return SERR_INVALID_STACK_OPERATION; dstack.pop_back(); return SERR_OK;
      case 0x77: if (dstack.size() < 2) return SERR_INVALID_STACK_OPERATION; dstack.erase(dstack.end() - 2); return SERR_OK;
      case 0x79: case 0x7a: {  // PICK ROLL
        if ((e = pop_nums(1, 4, n))) return e;
        if (n[0] < 0 || (size_t)n[0] >= dstack.size()) return op == 0x79 ? SERR_PICK_INVALID : SERR_ROLL_INVALID;
        size_t pos = dstack.size() - (size_t)n[0] - 1;
        Bytes item = dstack[pos];
        if (op == 0x7a) dstack.erase(dstack.begin() + pos);
        dstack.push_back(std::move(item));
        return SERR_OK;
      }
      case 0x7d: { if (dstack.size() < 2) return SERR_INVALID_STACK_OPERATION; Bytes top = dstack.back(); dstack.insert(dstack.end() - 2, std::move(top)); return SERR_OK; }
      case 0x7e: case 0x7f: case 0x80: case 0x81: case 0x83: case 0x84: case 0x85: case 0x86: case 0x8d: case 0x8e:
      case 0x95: case 0x96: case 0x97: case 0x98: case 0x99: return SERR_OPCODE_DISABLED;
      case 0x82: if (dstack.empty()) return SERR_INVALID_STACK_OPERATION; return push_num((int64_t)dstack.back().size());
      case 0x87: case 0x88: {
        if (dstack.size() < 2) return SERR_INVALID_STACK_OPERATION;
        bool eq = dstack[dstack.size() - 1] == dstack[dstack.size() - 2];
        dstack.resize(dstack.size() - 2);
        if (op == 0x87) { dstack.push_back(eq ? Bytes{1} : Bytes{}); return SERR_OK; }
        return eq ? SERR_OK : SERR_VERIFY;
      }
      case 0x8b: if ((e = pop_nums(1, 8, n))) return e; if (n[0] == INT64_MAX) return SERR_NUMBER_TOO_BIG; return push_num(n[0] + 1);
      case 0x8c: if ((e = pop_nums(1, 8, n))) return e; if (n[0] == INT64_MIN) return SERR_NUMBER_TOO_BIG; return push_num(n[0] - 1);
      case 0x8f: if ((e = pop_nums(1, 8, n))) return e; if (n[0] == INT64_MIN) return SERR_NUMBER_TOO_BIG; return push_num(-n[0]);
      case 0x90: if ((e = pop_nums(1, 8, n))) return e; if (n[0] == INT64_MIN) return SERR_NUMBER_TOO_BIG; return push_num(n[0] < 0 ? -n[0] : n[0]);
      case 0x91: if ((e = pop_nums(1, 8, n))) return e; return push_num(n[0] == 0);
      case 0x92: if ((e = pop_nums(1, 8, n))) return e; return push_num(n[0] != 0);
      case 0x93: { if ((e = pop_nums(2, 8, n))) return e; int64_t r; if (__builtin_add_overflow(n[0], n[1], &r)) return SERR_NUMBER_TOO_BIG; return push_num(r); }
      case 0x94: { if ((e = pop_nums(2, 8, n))) return e; int64_t r; if (__builtin_sub_overflow(n[0], n[1], &r)) return SERR_NUMBER_TOO_BIG; return push_num(r); }
      case 0x9a: if ((e = pop_nums(2, 8, n))) return e; return push_num(n[0] != 0 && n[1] != 0);
      case 0x9b: if ((e = pop_nums(2, 8, n))) return e; return push_num(n[0] != 0 || n[1] != 0);
      case 0x9c: if ((e = pop_nums(2, 8, n))) return e; return push_num(n[0] == n[1]);
      case 0x9d: if ((e = pop_nums(2, 8, n))) return e; return n[0] == n[1] ? SERR_OK : SERR_VERIFY;
      case 0x9e: if ((e = pop_nums(2, 8, n))) return e; return push_num(n[0] != n[1]);
      case 0x9f: if ((e = pop_nums(2, 8, n))) return e; return push_num(n[0] < n[1]);
      case 0xa0: if ((e = pop_nums(2, 8, n))) return e; return push_num(n[0] > n[1]);
      case 0xa1: if ((e = pop_nums(2, 8, n))) return e; return push_num(n[0] <= n[1]);
      case 0xa2: if ((e = pop_nums(2, 8, n))) return e; return push_num(n[0] >= n[1]);
      case 0xa3: if ((e = pop_nums(2, 8, n))) return e; return push_num(n[0] < n[1] ? n[0] : n[1]);
      case 0xa4: if ((e = pop_nums(2, 8, n))) return e; return push_num(n[0] > n[1] ? n[0] : n[1]);
      case 0xa5: if ((e = pop_nums(3, 8, n))) return e; return push_num(n[0] >= n[1] && n[0] < n[2]);
      case 0xa8: {  // OpSHA256
        if ((e = pop_raw(a))) return e;
        uint32_t st[8] = {0x6a09e667u, 0xbb67ae85u, 0x3c6ef372u, 0xa54ff53au, 0x510e527fu, 0x9b05688cu, 0x1f83d9abu, 0x5be0cd19u};
        Bytes m = a;
        uint64_t bits = (uint64_t)a.size() * 8;
        m.push_back(0x80);
        while (m.size() % 64 != 56) m.push_back(0);
        for (int i = 7; i >= 0; i--) m.push_back((uint8_t)(bits >> (8 * i)));
        for (size_t off = 0; off < m.size(); off += 64) {
          uint32_t w[16];
          for (int i = 0; i < 16; i++) w[i] = ((uint32_t)m[off + 4 * i] << 24) | ((uint32_t)m[off + 4 * i + 1] << 16) | ((uint32_t)m[off + 4 * i + 2] << 8) | m[off + 4 * i + 3];
          kgv::sha256_compress(st, w);
        }
        Bytes out(32);
        for (int i = 0; i < 8; i++) { out[4 * i] = (uint8_t)(st[i] >> 24); out[4 * i + 1] = (uint8_t)(st[i] >> 16); out[4 * i + 2] = (uint8_t)(st[i] >> 8); out[4 * i + 3] = (uint8_t)st[i]; }
        dstack.push_back(std::move(out));
        return SERR_OK;
      }
      case 0xaa: {  // OpBlake2b (unkeyed, 32 bytes)
        if ((e = pop_raw(a))) return e;
        kgv::Blake2b h;
        kgv::b2b_init(h, kgv::B2B_UNKEYED);
        kgv::b2b_bytes(h, a.data(), (uint32_t)a.size());
        uint64_t d[4];
        kgv::b2b_final(h, d);
        Bytes out(32);
        memcpy(out.data(), d, 32);
        dstack.push_back(std::move(out));
        return SERR_OK;
      }
      case 0xa9: return op_checkmultisig(true);
      case 0xab: return op_checksig(true);
      case 0xac: return op_checksig(false);
      case 0xad: { if ((e = op_checksig(false))) return e; bool v; if ((e = pop_bool(v))) return e; return v ? SERR_OK : SERR_VERIFY; }
      case 0xae: return op_checkmultisig(false);
      case 0xaf: { if ((e = op_checkmultisig(false))) return e; bool v; if ((e = pop_bool(v))) return e; return v ? SERR_OK : SERR_VERIFY; }
      case 0xb0: case 0xb1: {  // CLTV / CSV (opcodes/mod.rs:816-908)
        if ((e = pop_raw(a))) return e;
        if (a.size() > 8) return SERR_NUMBER_TOO_BIG;
        a.resize(8, 0);
        uint64_t v = 0;
        for (int i = 7; i >= 0; i--) v = (v << 8) | a[i];
        if (op == 0xb0) {
          bool both_lo = t.lock_time < LOCK_TIME_THRESHOLD && v < LOCK_TIME_THRESHOLD, both_hi = t.lock_time >= LOCK_TIME_THRESHOLD && v >= LOCK_TIME_THRESHOLD;
          if (!(both_lo || both_hi)) return SERR_UNSATISFIED_LOCKTIME;
          if (v > t.lock_time) return SERR_UNSATISFIED_LOCKTIME;
          if (in.sequence == UINT64_MAX) return SERR_UNSATISFIED_LOCKTIME;
          return SERR_OK;
        }
        if (v & SEQUENCE_LOCK_TIME_DISABLED) return SERR_OK;
        if (in.sequence & SEQUENCE_LOCK_TIME_DISABLED) return SERR_UNSATISFIED_LOCKTIME;
        if ((v & SEQUENCE_LOCK_TIME_MASK) > (in.sequence & SEQUENCE_LOCK_TIME_MASK)) return SERR_UNSATISFIED_LOCKTIME;
        return SERR_OK;
      }
      case 0xb2: case 0xb5: case 0xb6: case 0xb7: case 0xb8: case 0xba: case 0xbb: case 0xbc: case 0xbd: case 0xc0: case 0xc1: return SERR_OPCODE_RESERVED;
      case 0xb3: return push_num((int64_t)t.n_inputs);
      case 0xb4: return push_num((int64_t)t.n_outputs);
      case 0xb9: return push_num((int64_t)idx);
      case 0xbe: case 0xbf: {  // OpTxInputAmount / OpTxInputSpk
        if ((e = pop_nums(1, 4, n))) return e;
        if (n[0] < 0 || (uint64_t)n[0] >= t.n_inputs) return SERR_INVALID_INPUT_INDEX;
        const kgv_utxo_entry& u = b.entries[t.first_input + (uint32_t)n[0]];
        if (op == 0xbe) { if (u.amount > (uint64_t)INT64_MAX) return SERR_NUMBER_TOO_BIG; return push_num((int64_t)u.amount); }
        dstack.push_back(spk_to_bytes(u.spk_version, b.bytes + u.script_off, u.script_len));
        return SERR_OK;
      }
      case 0xc2: case 0xc3: {  // OpTxOutputAmount / OpTxOutputSpk
        if ((e = pop_nums(1, 4, n))) return e;
        if (n[0] < 0 || (uint64_t)n[0] >= t.n_outputs) return SERR_INVALID_OUTPUT_INDEX;
        const kgv_output& o = b.outputs[t.first_output + (uint32_t)n[0]];
        if (op == 0xc2) { if (o.value > (uint64_t)INT64_MAX) return SERR_NUMBER_TOO_BIG; return push_num((int64_t)o.value); }
        dstack.push_back(spk_to_bytes(o.spk_version, b.bytes + o.script_off, o.script_len));
        return SERR_OK;
      }
      default: return SERR_INVALID_OPCODE;  // 0xa6 0xa7 0xc4..0xff
    }
  }

  // execute_script (lib.rs:363-397)
  ScriptErr run_script(const uint8_t* s, size_t n, bool verify_only_push) {
    ScriptErr res = SERR_OK;
    size_t pos = 0;
    while (pos < n) {
      uint8_t op = s[pos++];
      const uint8_t* data = nullptr;
      size_t dlen = 0;
      if (op >= 0x01 && op <= 0x4b) {  // fixed-length pushes (macros.rs:27-41, :54-61)
        size_t avail = n - pos;
        if (avail < op) { res = SERR_MALFORMED_PUSH; break; }
        data = s + pos; dlen = op; pos += op;
      } else if (op >= 0x4c && op <= 0x4e) {  // macros.rs:9-26
        size_t lb = op == 0x4c ? 1 : op == 0x4d ? 2 : 4;
        if (n - pos < lb) { res = SERR_MALFORMED_PUSH_SIZE; break; }
        size_t l = 0;
        for (size_t i = 0; i < lb; i++) l |= (size_t)s[pos + i] << (8 * i);
        pos += lb;
        if (n - pos < l) { res = SERR_MALFORMED_PUSH; break; }
        data = s + pos; dlen = l; pos += l;
      }
      if (is_disabled(op)) { res = SERR_OPCODE_DISABLED; break; }
      if (op == 0x65 || op == 0x66) { res = SERR_OPCODE_RESERVED; break; }  // always illegal
      if (verify_only_push && op > NO_COST_OPCODE) { res = SERR_NOT_PUSH_ONLY; break; }
      // execute_opcode (lib.rs:322-344)
      if (op > NO_COST_OPCODE) {
        if (++num_ops > MAX_OPS_PER_SCRIPT) { res = SERR_TOO_MANY_OPERATIONS; break; }
      } else if (dlen > MAX_SCRIPT_ELEMENT_SIZE) { res = SERR_ELEMENT_TOO_BIG; break; }
      if (executing() || (op >= 0x63 && op <= 0x68)) {
        if (op > 0 && op <= 0x4e) { res = check_minimal_push(op, data, dlen); if (res) break; }
        res = exec(op, data, dlen);
        if (res) break;
      }
      if (astack.size() + dstack.size() > MAX_STACK_SIZE) { res = SERR_STACK_SIZE_EXCEEDED; break; }
    }
    if (res == SERR_OK && !cond.empty()) return SERR_UNBALANCED_CONDITIONAL;
    astack.clear();
    num_ops = 0;
    return res;
  }

  ScriptErr check_error_condition(bool final_script) {  // lib.rs:456-470
    if (final_script) {
      if (dstack.size() > 1) return SERR_CLEAN_STACK;
      if (dstack.empty()) return SERR_EMPTY_STACK;
    }
    bool v;
    ScriptErr e = pop_bool(v);
    if (e) return e;
    return v ? SERR_OK : SERR_EVAL_FALSE;
  }

  ScriptErr execute() {  // lib.rs:399-449
    if (entry.spk_version > 0) return SERR_OK;
    const uint8_t* ss = b.bytes + in.sigscript_off; size_t ssl = in.sigscript_len;
    const uint8_t* spk = b.bytes + entry.script_off; size_t spkl = entry.script_len;
    if (ssl == 0 && spkl == 0) return SERR_EVAL_FALSE;
    if (ssl > MAX_SCRIPTS_SIZE || spkl > MAX_SCRIPTS_SIZE) return SERR_SCRIPT_SIZE;
    bool p2sh = spkl == 35 && spk[0] == 0xaa && spk[1] == 0x20 && spk[34] == 0x87;
    Stack saved;
    bool have_saved = false;
    ScriptErr e;
    if (ssl) { if ((e = run_script(ss, ssl, true))) return e; }
    if (spkl) {
      if (p2sh) { saved = dstack; have_saved = true; }
      if ((e = run_script(spk, spkl, false))) return e;
    }
    if (p2sh) {
      if ((e = check_error_condition(false))) return e;
      if (!have_saved) return SERR_EMPTY_STACK;
      dstack = std::move(saved);
      if (dstack.empty()) return SERR_EMPTY_STACK;
      Bytes script = std::move(dstack.back()); dstack.pop_back();
      if ((e = run_script(script.data(), script.size(), false))) return e;
    }
    return check_error_condition(true);
  }
};

ScriptErr execute_input(const kgv_tx_batch& b, uint32_t tx, uint32_t input_index, const VerdictFn& verdict, std::vector<SigRequest>* missing) {
  Engine eng(b, tx, input_index, verdict, missing);
  return eng.execute();
}

}  // namespace kgv_host
