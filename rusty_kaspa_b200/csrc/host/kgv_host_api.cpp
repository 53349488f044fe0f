// kgv_host_api.cpp — C ABI of the host script engine (include/kgv.h): kgv_script_execute, kgv_check_scripts_host.
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "kgv_script_vm.h"
#include "../kgv_internal.h"  // kgv_ctx::err

using namespace kgv_host;

static void to_c_request(kgv_sig_request& o, const SigRequest& r) {
  memset(&o, 0, sizeof o);
  o.tx = r.tx; o.input = r.input_abs; o.hash_type = r.hash_type; o.ecdsa = r.ecdsa; o.key_len = (uint8_t)r.key.size();
  memcpy(o.key, r.key.data(), r.key.size() <= 33 ? r.key.size() : 33);
  memcpy(o.sig, r.sig, 64);
}

extern "C" int kgv_script_execute(const kgv_tx_batch* batch, uint32_t tx, uint32_t input_index, kgv_verdict_fn verdict, void* user, uint8_t* script_err) {
  if (!batch || !script_err || !batch->txs || tx >= batch->n_txs || input_index >= batch->txs[tx].n_inputs || !batch->entries) return KGV_ERR_ARG;
  VerdictFn fn = [&](const SigRequest& r) -> int {
    if (!verdict) return -1;
    kgv_sig_request c;
    to_c_request(c, r);
    return verdict(user, &c);
  };
  *script_err = (uint8_t)execute_input(*batch, tx, input_index, fn, nullptr);
  return KGV_OK;
}

static std::string request_key(const SigRequest& r) {
  std::string k;
  k.reserve(10 + r.key.size() + 64);
  k.append((const char*)&r.input_abs, 4);
  k.push_back((char)r.hash_type);
  k.push_back((char)r.ecdsa);
  k.append((const char*)r.key.data(), r.key.size());
  k.append((const char*)r.sig, 64);
  return k;
}

extern "C" int kgv_check_scripts_host(kgv_ctx* ctx, const kgv_tx_batch* batch, const uint32_t* tx_indices, size_t n, kgv_tx_result* results) {
  if (!ctx || !batch || (n && (!tx_indices || !results)) || !batch->entries) return KGV_ERR_ARG;
  std::map<std::string, int> verdicts;
  std::vector<char> done(n, 0);
  for (size_t i = 0; i < n; i++) {
    if (tx_indices[i] >= batch->n_txs) return KGV_ERR_ARG;
    memset(&results[i], 0, sizeof results[i]);
  }
  for (int round = 0; round < 4096; round++) {
    std::vector<SigRequest> missing;
    VerdictFn fn = [&](const SigRequest& r) -> int {
      auto it = verdicts.find(request_key(r));
      return it == verdicts.end() ? -1 : it->second;
    };
    bool pending = false;
    for (size_t i = 0; i < n; i++) {
      if (done[i]) continue;
      uint32_t ti = tx_indices[i];
      const kgv_tx& t = batch->txs[ti];
      bool suspended = false;
      kgv_tx_result r;
      memset(&r, 0, sizeof r);
      for (uint32_t k = 0; k < t.n_inputs; k++) {  // check_scripts_sequential order
        ScriptErr e = execute_input(*batch, ti, k, fn, &missing);
        if (e == SERR_NEEDS_SIG_VERDICTS) { suspended = true; continue; }  // keep going: the later inputs' requests join the same GPU round
        if (suspended) continue;                                            // (their outcomes only count once every earlier input is decided)
        if (e != SERR_OK) {
          r.fail_input = k;
          r.script_err = (uint8_t)e;
          r.status = batch->inputs[t.first_input + k].sigscript_len == 0 ? KGV_TX_SIGNATURE_EMPTY : KGV_TX_SIGNATURE_INVALID;
          break;
        }
      }
      if (suspended) { pending = true; continue; }
      results[i] = r;
      done[i] = 1;
    }
    if (!pending) return KGV_OK;
    // verify everything that was asked for, on the GPU: sighash -> Schnorr / ECDSA batches
    std::vector<SigRequest> uniq;
    for (auto& m : missing)
      if (!verdicts.count(request_key(m))) { verdicts[request_key(m)] = -2; uniq.push_back(m); }
    for (int ecdsa = 0; ecdsa < 2; ecdsa++) {
      std::vector<const SigRequest*> sel;
      for (auto& u : uniq) if (u.ecdsa == ecdsa) sel.push_back(&u);
      if (sel.empty()) continue;
      size_t m = sel.size(), kl = ecdsa ? 33 : 32;
      std::vector<kgv_sighash_item> items(m);
      std::vector<uint8_t> pk(m * kl), sig(m * 64), msg(m * 32), st(m);
      for (size_t j = 0; j < m; j++) {
        memset(&items[j], 0, sizeof items[j]);
        items[j].tx = sel[j]->tx; items[j].input = sel[j]->input_abs; items[j].hash_type = sel[j]->hash_type; items[j].ecdsa = (uint8_t)ecdsa;
        memcpy(&pk[j * kl], sel[j]->key.data(), kl);
        memcpy(&sig[j * 64], sel[j]->sig, 64);
      }
      int rc = kgv_sighash(ctx, batch, items.data(), m, msg.data());
      if (rc) return rc;
      rc = ecdsa ? kgv_ecdsa_verify(ctx, pk.data(), msg.data(), sig.data(), m, st.data()) : kgv_schnorr_verify(ctx, pk.data(), msg.data(), sig.data(), m, st.data());
      if (rc) return rc;
      for (size_t j = 0; j < m; j++) verdicts[request_key(*sel[j])] = st[j];
    }
  }
  // one round resolves at least one pending check per script, and a script can reach at most 201 + 20 x 201 checks (ops limit, multisig keys):
  // only scripts beyond the engine's own limits could get here
  ctx->err = "kgv_check_scripts_host: a script still asks for signature verdicts after 4096 verification rounds";
  return KGV_ERR_LIMIT;
}

// ---------------------------------------------------------------------------------------------
// Persistence formats either side of the path: the rows DbUtxoSetStore keeps in RocksDB.
//   key   consensus/src/model/stores/utxo_set.rs:31-88   UtxoKey = txid(32) || index u32 LE, stored with the trailing zero bytes of the index
//         trimmed but at least one index byte kept (AsRef<[u8]> :36-44); TryFrom pads a short slice back with zeros (:46-62)
//   value database/src/access.rs:139 bincode::serialize(&UtxoEntry) with bincode 1.x defaults (fixed-width little-endian integers, u64 lengths):
//         amount u64 || script_public_key.version u16 || script length u64 || script bytes || block_daa_score u64 || is_coinbase u8
//         (UtxoEntry consensus/core/src/tx.rs:49-57; ScriptPublicKeyInternal {version, script: &[u8]} tx/script_public_key.rs:73-94)
// Host code by design: RocksDB lives on the host; these are the (de)serialisers a shim uses between the store and kgv_utxo_apply_diff /
// kgv_utxo_lookup.  Batch forms: rows are written back to back, row_off[i] .. row_off[i+1].
// ---------------------------------------------------------------------------------------------
extern "C" int kgv_utxo_rows_encode(const uint8_t* keys36, const kgv_utxo_entry* entries, const uint8_t* bytes, size_t n_bytes, size_t n, uint8_t* key_rows, uint64_t* key_off,
                                    uint8_t* value_rows, uint64_t* value_off, size_t key_cap, size_t value_cap) {
  if ((n && (!keys36 || !entries || !key_off || !value_off)) || (n && !key_rows && key_cap) || (n && !value_rows && value_cap)) return KGV_ERR_ARG;
  uint64_t ko = 0, vo = 0;
  for (size_t i = 0; i < n; i++) {
    const uint8_t* k = keys36 + 36 * i;
    int last = 0;  // rposition of a non-zero index byte, 0 if none
    for (int b = 3; b >= 0; b--) if (k[32 + b]) { last = b; break; }
    const size_t klen = 32 + (size_t)last + 1;
    const kgv_utxo_entry& e = entries[i];
    if ((uint64_t)e.script_off + e.script_len > n_bytes) return KGV_ERR_ARG;
    const size_t vlen = 8 + 2 + 8 + (size_t)e.script_len + 8 + 1;
    key_off[i] = ko; value_off[i] = vo;
    if (key_rows) { if (ko + klen > key_cap) return KGV_ERR_NOMEM; memcpy(key_rows + ko, k, klen); }
    if (value_rows) {
      if (vo + vlen > value_cap) return KGV_ERR_NOMEM;
      uint8_t* p = value_rows + vo;
      for (int b = 0; b < 8; b++) *p++ = (uint8_t)(e.amount >> (8 * b));
      *p++ = (uint8_t)e.spk_version; *p++ = (uint8_t)(e.spk_version >> 8);
      for (int b = 0; b < 8; b++) *p++ = (uint8_t)((uint64_t)e.script_len >> (8 * b));
      memcpy(p, bytes + e.script_off, e.script_len); p += e.script_len;
      for (int b = 0; b < 8; b++) *p++ = (uint8_t)(e.block_daa_score >> (8 * b));
      *p++ = e.is_coinbase ? 1 : 0;
    }
    ko += klen; vo += vlen;
  }
  if (n) { key_off[n] = ko; value_off[n] = vo; }
  return KGV_OK;
}

extern "C" int kgv_utxo_rows_decode(const uint8_t* key_rows, const uint64_t* key_off, const uint8_t* value_rows, const uint64_t* value_off, size_t n, uint8_t* keys36,
                                    kgv_utxo_entry* entries, uint8_t* bytes_out, size_t bytes_cap, size_t* bytes_used) {
  if (n && (!key_rows || !key_off || !value_rows || !value_off || !keys36 || !entries)) return KGV_ERR_ARG;
  size_t used = 0;
  for (size_t i = 0; i < n; i++) {
    const size_t klen = (size_t)(key_off[i + 1] - key_off[i]), vlen = (size_t)(value_off[i + 1] - value_off[i]);
    if (klen < 33 || klen > 36) return KGV_ERR_ARG;  // "src slice is too short" / "too large" (utxo_set.rs:49-55)
    memset(keys36 + 36 * i, 0, 36);
    memcpy(keys36 + 36 * i, key_rows + key_off[i], klen);
    const uint8_t* p = value_rows + value_off[i];
    if (vlen < 27) return KGV_ERR_ARG;
    kgv_utxo_entry e;
    memset(&e, 0, sizeof e);
    for (int b = 0; b < 8; b++) e.amount |= (uint64_t)p[b] << (8 * b);
    e.spk_version = (uint16_t)(p[8] | (p[9] << 8));
    uint64_t sl = 0;
    for (int b = 0; b < 8; b++) sl |= (uint64_t)p[10 + b] << (8 * b);
    if (sl != vlen - 27) return KGV_ERR_ARG;  // bincode would report a length mismatch / trailing bytes
    if (used + sl > bytes_cap || used + sl > 0xFFFFFFFFull) return KGV_ERR_NOMEM;
    if (sl) memcpy(bytes_out + used, p + 18, sl);
    e.script_off = (uint32_t)used; e.script_len = (uint32_t)sl;
    const uint8_t* q = p + 18 + sl;
    for (int b = 0; b < 8; b++) e.block_daa_score |= (uint64_t)q[b] << (8 * b);
    if (q[8] > 1) return KGV_ERR_ARG;         // bincode: invalid bool encoding
    e.is_coinbase = q[8];
    entries[i] = e;
    used += sl;
  }
  if (bytes_used) *bytes_used = used;
  return KGV_OK;
}
