// kgv_host_api.cpp — C ABI of the host script engine (include/kgv.h): kgv_script_execute, kgv_check_scripts_host.
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "kgv_script_vm.h"

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
        if (e == SERR_NEEDS_SIG_VERDICTS) { suspended = true; break; }
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
  return KGV_ERR_ARG;  // unreachable for scripts within the op-count limits
}
