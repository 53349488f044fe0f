/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the shipped product path.
 *
 * Transaction id / hash and signature-hash restatement:
 *   consensus/core/src/hashing/tx.rs:16-107      (write_transaction, flags FULL / EXCLUDE_*)
 *   consensus/core/src/hashing/sighash.rs:140-277 (sub-hashes, calc_schnorr/ecdsa_signature_hash)
 *   consensus/core/src/hashing/mod.rs:46-96       (write_len = u64 LE, write_var_bytes)
 * Pinned by the reference's vectors in tests/golden/{tx_hashing,sighash}.json.
 */
#include "ok_oracle.h"
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

static void w_u8(ok_blake2b_ctx* h, uint8_t v) { ok_blake2b_update(h, &v, 1); }
static void w_u16(ok_blake2b_ctx* h, uint16_t v) { uint8_t b[2] = {(uint8_t)v, (uint8_t)(v >> 8)}; ok_blake2b_update(h, b, 2); }
static void w_u32(ok_blake2b_ctx* h, uint32_t v) { uint8_t b[4]; for (int i = 0; i < 4; i++) b[i] = (uint8_t)(v >> (8 * i)); ok_blake2b_update(h, b, 4); }
static void w_u64(ok_blake2b_ctx* h, uint64_t v) { uint8_t b[8]; for (int i = 0; i < 8; i++) b[i] = (uint8_t)(v >> (8 * i)); ok_blake2b_update(h, b, 8); }
static void w_var(ok_blake2b_ctx* h, const uint8_t* p, size_t n) { w_u64(h, (uint64_t)n); if (n) ok_blake2b_update(h, p, n); }
static void hinit(ok_blake2b_ctx* h, const char* domain) { ok_blake2b_init(h, 32, domain, strlen(domain)); }

static int is_coinbase(const ok_tx* t) { /* subnets::SUBNETWORK_ID_COINBASE = 01 00..00 */
  if (t->subnetwork_id[0] != 1) return 0;
  for (int i = 1; i < 20; i++) if (t->subnetwork_id[i]) return 0;
  return 1;
}
static int is_native(const ok_tx* t) {
  for (int i = 0; i < 20; i++) if (t->subnetwork_id[i]) return 0;
  return 1;
}

/* hashing/tx.rs:45-107 */
static void write_tx(ok_blake2b_ctx* h, const ok_batch* b, const ok_tx* t, int exclude_sigscript, int exclude_mass) {
  w_u16(h, t->version);
  w_u64(h, t->n_inputs);
  for (uint32_t i = 0; i < t->n_inputs; i++) {
    const ok_input* in = &b->inputs[t->first_input + i];
    ok_blake2b_update(h, in->prev_txid, 32);
    w_u32(h, in->prev_index);
    if (!exclude_sigscript) {
      w_var(h, b->bytes + in->sigscript_off, in->sigscript_len);
      w_u8(h, in->sig_op_count);
    } else {
      w_var(h, NULL, 0);
    }
    w_u64(h, in->sequence);
  }
  w_u64(h, t->n_outputs);
  for (uint32_t i = 0; i < t->n_outputs; i++) {
    const ok_output* o = &b->outputs[t->first_output + i];
    w_u64(h, o->value);
    w_u16(h, o->spk_version);
    w_var(h, b->bytes + o->script_off, o->script_len);
  }
  w_u64(h, t->lock_time);
  ok_blake2b_update(h, t->subnetwork_id, 20);
  w_u64(h, t->gas);
  w_var(h, b->bytes + t->payload_off, t->payload_len);
  if (!exclude_mass && t->mass > 0) w_u64(h, t->mass);
}

void ok_tx_id(const ok_batch* b, size_t tx, uint8_t out[32]) {
  const ok_tx* t = &b->txs[tx];
  int cb = is_coinbase(t);
  ok_blake2b_ctx h;
  hinit(&h, "TransactionID");
  write_tx(&h, b, t, !cb, !cb);
  ok_blake2b_final(&h, out);
}
void ok_tx_hash(const ok_batch* b, size_t tx, uint8_t out[32]) {
  ok_blake2b_ctx h;
  hinit(&h, "TransactionHash");
  write_tx(&h, b, &b->txs[tx], 0, 0);
  ok_blake2b_final(&h, out);
}

typedef struct { const ok_batch* b; uint8_t* out; size_t lo, hi; int hash; } hjob;
static void* hworker(void* a) {
  hjob* j = (hjob*)a;
  for (size_t i = j->lo; i < j->hi; i++) { if (j->hash) ok_tx_hash(j->b, i, j->out + 32 * i); else ok_tx_id(j->b, i, j->out + 32 * i); }
  return NULL;
}
static void hrun(const ok_batch* b, uint8_t* out, int nthreads, int hash) {
  size_t n = b->n_txs;
  if (nthreads < 1) nthreads = 1;
  if ((size_t)nthreads > n) nthreads = n ? (int)n : 1;
  pthread_t* th = malloc(sizeof(pthread_t) * nthreads);
  hjob* jobs = malloc(sizeof(hjob) * nthreads);
  for (int t = 0; t < nthreads; t++) {
    jobs[t] = (hjob){b, out, n * t / nthreads, n * (t + 1) / nthreads, hash};
    if (t + 1 < nthreads) pthread_create(&th[t], NULL, hworker, &jobs[t]);
  }
  hworker(&jobs[nthreads - 1]);
  for (int t = 0; t + 1 < nthreads; t++) pthread_join(th[t], NULL);
  free(th); free(jobs);
}
void ok_tx_ids(const ok_batch* b, uint8_t* out32, int nthreads) { hrun(b, out32, nthreads, 0); }
void ok_tx_hashes(const ok_batch* b, uint8_t* out32, int nthreads) { hrun(b, out32, nthreads, 1); }

/* ---------------------------------------------------------------- sighash */
#define SH_ALL 1
#define SH_NONE 2
#define SH_SINGLE 4
#define SH_ACP 0x80
/* sighash_type.rs:24-48: the low bits are compared after masking with 0b111 */
static int ht_single(uint8_t t) { return (t & 7) == SH_SINGLE; }
static int ht_none(uint8_t t) { return (t & 7) == SH_NONE; }
static int ht_acp(uint8_t t) { return (t & SH_ACP) != 0; }

static void hash_output(ok_blake2b_ctx* h, const ok_batch* b, const ok_output* o) {
  w_u64(h, o->value);
  w_u16(h, o->spk_version);
  w_var(h, b->bytes + o->script_off, o->script_len);
}

void ok_sighash(const ok_batch* b, const ok_utxo_entry* entries, size_t tx, uint32_t input_index, uint8_t hash_type, int ecdsa, uint8_t out[32]) {
  const ok_tx* t = &b->txs[tx];
  uint8_t prev[32] = {0}, seqs[32] = {0}, sops[32] = {0}, outs[32] = {0}, pay[32] = {0};
  ok_blake2b_ctx h;
  if (!ht_acp(hash_type)) { /* sighash.rs:140-153 */
    hinit(&h, "TransactionSigningHash");
    for (uint32_t i = 0; i < t->n_inputs; i++) { const ok_input* in = &b->inputs[t->first_input + i]; ok_blake2b_update(&h, in->prev_txid, 32); w_u32(&h, in->prev_index); }
    ok_blake2b_final(&h, prev);
  }
  if (!(ht_single(hash_type) || ht_acp(hash_type) || ht_none(hash_type))) { /* :155-167 */
    hinit(&h, "TransactionSigningHash");
    for (uint32_t i = 0; i < t->n_inputs; i++) w_u64(&h, b->inputs[t->first_input + i].sequence);
    ok_blake2b_final(&h, seqs);
  }
  if (!ht_acp(hash_type)) { /* :169-182 */
    hinit(&h, "TransactionSigningHash");
    for (uint32_t i = 0; i < t->n_inputs; i++) w_u8(&h, b->inputs[t->first_input + i].sig_op_count);
    ok_blake2b_final(&h, sops);
  }
  if (!(is_native(t) && t->payload_len == 0)) { /* :184-195 */
    hinit(&h, "TransactionSigningHash");
    w_var(&h, b->bytes + t->payload_off, t->payload_len);
    ok_blake2b_final(&h, pay);
  }
  if (ht_none(hash_type)) { /* :197-221 */
  } else if (ht_single(hash_type)) {
    if (input_index < t->n_outputs) {
      hinit(&h, "TransactionSigningHash");
      hash_output(&h, b, &b->outputs[t->first_output + input_index]);
      ok_blake2b_final(&h, outs);
    }
  } else {
    hinit(&h, "TransactionSigningHash");
    for (uint32_t i = 0; i < t->n_outputs; i++) hash_output(&h, b, &b->outputs[t->first_output + i]);
    ok_blake2b_final(&h, outs);
  }
  const ok_input* in = &b->inputs[t->first_input + input_index];
  const ok_utxo_entry* e = &entries[t->first_input + input_index];
  hinit(&h, "TransactionSigningHash"); /* :238-265 */
  w_u16(&h, t->version);
  ok_blake2b_update(&h, prev, 32);
  ok_blake2b_update(&h, seqs, 32);
  ok_blake2b_update(&h, sops, 32);
  ok_blake2b_update(&h, in->prev_txid, 32);
  w_u32(&h, in->prev_index);
  w_u16(&h, e->spk_version);
  w_var(&h, b->bytes + e->script_off, e->script_len);
  w_u64(&h, e->amount);
  w_u64(&h, in->sequence);
  w_u8(&h, in->sig_op_count);
  ok_blake2b_update(&h, outs, 32);
  w_u64(&h, t->lock_time);
  ok_blake2b_update(&h, t->subnetwork_id, 20);
  w_u64(&h, t->gas);
  ok_blake2b_update(&h, pay, 32);
  w_u8(&h, hash_type);
  uint8_t sh[32];
  ok_blake2b_final(&h, sh);
  if (ecdsa) ok_sha256_domain("TransactionSigningHashECDSA", sh, 32, out); /* :267-277 */
  else memcpy(out, sh, 32);
}

/* crypto/merkle/src/lib.rs:3-30 calc_merkle_root + merkle_hash (keyed BLAKE2b "MerkleBranchHash" of left || right):
 * levels are padded to the next power of two; a pair whose left child is missing is missing, a missing right child
 * counts as ZERO_HASH; no hashes at all -> ZERO_HASH.  Used for hash_merkle_root (consensus/core/src/merkle.rs:5-7,
 * body_validation_in_isolation.rs:34-40) and accepted_id_merkle_root (utxo_validation.rs:401-410). */
void ok_merkle_root(const uint8_t* hashes32, size_t n, uint8_t out[32]) {
  if (n == 0) { memset(out, 0, 32); return; }
  size_t pot = 1;
  while (pot < n) pot <<= 1;
  uint8_t* cur = (uint8_t*)malloc(pot * 32);
  memcpy(cur, hashes32, n * 32);
  size_t have = n; /* entries [0, have) present, the rest missing */
  for (size_t width = pot; width > 1; width >>= 1) {
    size_t nh = (have + 1) / 2;
    for (size_t i = 0; i < nh; i++) {
      uint8_t buf[64];
      memcpy(buf, cur + 64 * i, 32);
      if (2 * i + 1 < have) memcpy(buf + 32, cur + 64 * i + 32, 32); else memset(buf + 32, 0, 32);
      ok_blake2b_keyed("MerkleBranchHash", buf, 64, cur + 32 * i);
    }
    have = nh;
  }
  memcpy(out, cur, 32);
  free(cur);
}

/* body_validation_in_isolation.rs:13-23,95-131: check_duplicate_transactions, check_block_double_spends,
 * check_no_chained_transactions for the block made of transactions [t0, t1) of the batch; returns 0 ok, 1 duplicate
 * transactions, 2 double spend in the same block, 3 chained transaction; *index = first offender in the reference's
 * iteration order (tx index / absolute input index). */
int ok_block_set_checks(const ok_batch* b, uint32_t t0, uint32_t t1, uint32_t* index) {
  *index = 0;
  if (t0 == t1) return 0;
  uint32_t nt = t1 - t0;
  uint8_t* ids = (uint8_t*)malloc((size_t)nt * 32);
  for (uint32_t t = 0; t < nt; t++) ok_tx_id(b, t0 + t, ids + 32 * (size_t)t);
  int rc = 0;
  for (uint32_t t = 0; t < nt && !rc; t++)
    for (uint32_t j = 0; j < t; j++)
      if (!memcmp(ids + 32 * (size_t)j, ids + 32 * (size_t)t, 32)) { rc = 1; *index = t0 + t; break; }
  uint32_t i0 = b->txs[t0].first_input, i1 = b->txs[t1 - 1].first_input + b->txs[t1 - 1].n_inputs;
  for (uint32_t i = i0; i < i1 && !rc; i++)
    for (uint32_t j = i0; j < i; j++)
      if (b->inputs[j].prev_index == b->inputs[i].prev_index && !memcmp(b->inputs[j].prev_txid, b->inputs[i].prev_txid, 32)) { rc = 2; *index = i; break; }
  for (uint32_t i = i0; i < i1 && !rc; i++)
    for (uint32_t t = 0; t < nt; t++)
      if (b->inputs[i].prev_index < b->txs[t0 + t].n_outputs && !memcmp(ids + 32 * (size_t)t, b->inputs[i].prev_txid, 32)) { rc = 3; *index = i; break; }
  free(ids);
  return rc;
}
