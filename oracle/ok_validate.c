/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the shipped product path.
 *
 * UTXO-context transaction validation, restated from
 *   consensus/src/processes/transaction_validator/tx_validation_in_utxo_context.rs:34-200
 *   consensus/core/src/mass/mod.rs:64-80,338-410                      (storage mass, KIP-9)
 *   consensus/core/src/utxo/{utxo_view.rs:5-50, utxo_diff.rs:233-269}  (composed view, diff rules)
 *   consensus/src/model/stores/utxo_set.rs:107-112                     (write_diff_batch)
 *   consensus/src/pipeline/virtual_processor/utxo_validation.rs:262-338
 *   crypto/txscript/src/lib.rs:399-643 for the three standard script shapes only.
 */
#include "ok_oracle.h"
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

static int g_fast_verify = 0;
void ok_use_fast_verify(int on) { g_fast_verify = on; }

/* ------------------------------------------------------------------ standard-class script check */
static int sighash_type_ok(uint8_t t) { return t == 1 || t == 2 || t == 4 || t == 0x81 || t == 0x82 || t == 0x84; }

/* lib.rs:574-608 / :610-643 — returns OK_SCRIPT_* error (<0 none): *valid set on success */
static int check_sig(const ok_batch* b, const ok_utxo_entry* entries, size_t tx, uint32_t idx, int* remaining, uint8_t hash_type, const uint8_t* key, size_t keylen,
                     const uint8_t* sig, size_t siglen, int ecdsa, int* valid) {
  if (*remaining == 0) return OK_SCRIPT_EXCEEDED_SIGOP_LIMIT; /* runtime_sig_op_counter.rs:40-44, before any length check */
  (*remaining)--;
  if (siglen != 64) return OK_SCRIPT_SIG_LENGTH;
  if (keylen != (size_t)(ecdsa ? 33 : 32)) return OK_SCRIPT_PUBKEY_FORMAT;
  uint8_t msg[32];
  ok_sighash(b, entries, tx, idx, hash_type, ecdsa, msg);
  int st = g_fast_verify ? (ecdsa ? ok_ecdsa_verify_fast(key, msg, sig) : ok_schnorr_verify_fast(key, msg, sig))
                         : (ecdsa ? ok_ecdsa_verify(key, msg, sig) : ok_schnorr_verify(key, msg, sig));
  if (st == OK_SIG_PK_PARSE_ERR || st == OK_SIG_SIG_PARSE_ERR) return OK_SCRIPT_INVALID_SIGNATURE;
  *valid = st == OK_SIG_VALID;
  return -1;
}

/* parse one canonical data push at p (lib.rs opcode parser + minimal push rule, opcodes/mod.rs:141-190);
 * returns consumed bytes or 0 if not a canonical direct push */
static size_t canonical_push(const uint8_t* p, size_t n, const uint8_t** data, size_t* len) {
  if (n == 0) return 0;
  uint8_t op = p[0];
  if (op >= 1 && op <= 75) {
    if (n < 1u + op) return 0;
    if (op == 1 && ((p[1] >= 1 && p[1] <= 16) || p[1] == 0x81)) return 0; /* must have used OP_1..16 / OP_1NEGATE */
    *data = p + 1; *len = op;
    return 1u + op;
  }
  if (op == 0x4c) {
    if (n < 2) return 0;
    size_t l = p[1];
    if (l <= 75 || n < 2 + l) return 0;
    *data = p + 2; *len = l;
    return 2 + l;
  }
  if (op == 0x4d) {
    if (n < 3) return 0;
    size_t l = (size_t)p[1] | ((size_t)p[2] << 8);
    if (l <= 255 || n < 3 + l) return 0;
    *data = p + 3; *len = l;
    return 3 + l;
  }
  return 0;
}

int ok_check_script_std(const ok_batch* b, const ok_utxo_entry* entries, size_t tx, uint32_t input_index) {
  const ok_tx* t = &b->txs[tx];
  const ok_input* in = &b->inputs[t->first_input + input_index];
  const ok_utxo_entry* e = &entries[t->first_input + input_index];
  const uint8_t* spk = b->bytes + e->script_off;
  size_t spk_len = e->script_len;
  const uint8_t* ss = b->bytes + in->sigscript_off;
  size_t ss_len = in->sigscript_len;
  if (e->spk_version > 0) return OK_SCRIPT_OK; /* lib.rs:402-405 */
  int remaining = in->sig_op_count;
  int ecdsa;
  if ((spk_len == 34 && spk[0] == 0x20 && spk[33] == 0xac) || (spk_len == 35 && spk[0] == 0x21 && spk[34] == 0xab)) {
    /* P2PK (script_class.rs:58-70): sigscript must be exactly one 65-byte push */
    ecdsa = spk_len == 35;
    if (!(ss_len == 66 && ss[0] == 0x41)) return OK_SCRIPT_NONSTANDARD;
    uint8_t ht = ss[65];
    if (!sighash_type_ok(ht)) return OK_SCRIPT_INVALID_SIGHASH_TYPE; /* opcodes/mod.rs:751,774 */
    int valid = 0;
    int err = check_sig(b, entries, tx, input_index, &remaining, ht, spk + 1, spk_len - 2, ss + 1, 64, ecdsa, &valid);
    if (err >= 0) return err;
    return valid ? OK_SCRIPT_OK : OK_SCRIPT_EVAL_FALSE; /* lib.rs:456-470 */
  }
  if (spk_len == 35 && spk[0] == 0xaa && spk[1] == 0x20 && spk[34] == 0x87) {
    /* P2SH (script_class.rs:72-82) over a standard multisig redeem script (standard/multisig.rs:18-70) */
    const uint8_t* sigs[20];
    size_t nsig = 0, off = 0;
    const uint8_t *redeem = NULL; size_t rlen = 0;
    while (off < ss_len) {
      const uint8_t* d; size_t l;
      size_t used = canonical_push(ss + off, ss_len - off, &d, &l);
      if (!used) return OK_SCRIPT_NONSTANDARD;
      off += used;
      if (off == ss_len) { redeem = d; rlen = l; break; }
      if (l != 65 || nsig == 20) return OK_SCRIPT_NONSTANDARD;
      sigs[nsig++] = d;
    }
    if (!redeem || rlen < 3 || rlen > 520) return OK_SCRIPT_NONSTANDARD;
    uint8_t last = redeem[rlen - 1];
    if (last != 0xae && last != 0xa9) return OK_SCRIPT_NONSTANDARD;
    ecdsa = last == 0xa9;
    if (redeem[0] < 0x51 || redeem[0] > 0x60) return OK_SCRIPT_NONSTANDARD;
    size_t m = redeem[0] - 0x50, klen = ecdsa ? 33 : 32;
    const uint8_t* keys[20];
    size_t nkeys = 0, p = 1;
    while (p < rlen - 2) {
      if (redeem[p] != klen || p + 1 + klen > rlen - 2 || nkeys == 20) return OK_SCRIPT_NONSTANDARD;
      keys[nkeys++] = redeem + p + 1;
      p += 1 + klen;
    }
    if (p != rlen - 2) return OK_SCRIPT_NONSTANDARD;
    uint8_t opn = redeem[rlen - 2];
    if (opn < 0x51 || opn > 0x60 || (size_t)(opn - 0x50) != nkeys || m > nkeys || m != nsig) return OK_SCRIPT_NONSTANDARD;
    /* spk: OpBlake2b OpData32 OpEqual, then check_error_condition(false) (lib.rs:441-442) */
    uint8_t h[32];
    ok_blake2b_256(redeem, rlen, h);
    if (memcmp(h, spk + 2, 32) != 0) return OK_SCRIPT_EVAL_FALSE;
    /* op_check_multisig_schnorr_or_ecdsa (lib.rs:488-571); sigs are all 65 bytes (non-empty) here */
    int failed = 0;
    size_t ki = 0;
    for (size_t si = 0; si < nsig && !failed; si++) {
      uint8_t ht = sigs[si][64];
      if (!sighash_type_ok(ht)) return OK_SCRIPT_INVALID_SIGHASH_TYPE;
      for (;;) {
        if (nkeys - ki < nsig - si) { failed = 1; break; }
        const uint8_t* key = keys[ki++];
        int valid = 0;
        int err = check_sig(b, entries, tx, input_index, &remaining, ht, key, klen, sigs[si], 64, ecdsa, &valid);
        if (err >= 0) return err;
        if (valid) break;
      }
    }
    if (failed) return OK_SCRIPT_NULL_FAIL; /* every signature here is non-empty (lib.rs:564-566) */
    return OK_SCRIPT_OK;
  }
  return OK_SCRIPT_NONSTANDARD;
}

/* ------------------------------------------------------------------ storage mass (mass/mod.rs) */
static uint64_t plurality(uint32_t script_len) { return (63u + (uint64_t)script_len + 99u) / 100u; } /* :64-80 */
static int ckmul(uint64_t a, uint64_t b, uint64_t* r) { return __builtin_mul_overflow(a, b, r); }
static int ckadd(uint64_t a, uint64_t b, uint64_t* r) { return __builtin_add_overflow(a, b, r); }
static uint64_t satadd(uint64_t a, uint64_t b) { uint64_t r; return ckadd(a, b, &r) ? UINT64_MAX : r; }
static uint64_t satsub(uint64_t a, uint64_t b) { return a > b ? a - b : 0; }
static uint64_t satmul(uint64_t a, uint64_t b) { uint64_t r; return ckmul(a, b, &r) ? UINT64_MAX : r; }

static int tx_is_coinbase(const ok_tx* t) {
  if (t->subnetwork_id[0] != 1) return 0;
  for (int i = 1; i < 20; i++) if (t->subnetwork_id[i]) return 0;
  return 1;
}

int ok_storage_mass(const ok_batch* b, const ok_utxo_entry* entries, size_t tx, uint64_t C, uint64_t* mass) {
  const ok_tx* t = &b->txs[tx];
  if (tx_is_coinbase(t)) { *mass = 0; return 0; }
  uint64_t outs_plur = 0, harm_outs = 0;
  for (uint32_t i = 0; i < t->n_outputs; i++) { /* :363-373 */
    const ok_output* o = &b->outputs[t->first_output + i];
    uint64_t p = plurality(o->script_len), v;
    outs_plur += p;
    if (ckmul(C, p, &v) || ckmul(v, p, &v)) return -1;
    if (o->value == 0) return -1; /* the reference would divide by zero: ruled out earlier by in-isolation checks */
    if (ckadd(harm_outs, v / o->value, &harm_outs)) return -1;
  }
  int relaxed;
  if (outs_plur == 1) relaxed = 1;
  else if (t->n_inputs > 2) relaxed = 0;
  else {
    uint64_t ip = 0;
    for (uint32_t i = 0; i < t->n_inputs; i++) ip += plurality(entries[t->first_input + i].script_len);
    relaxed = ip == 1 || (outs_plur == 2 && ip == 2);
  }
  if (relaxed) { /* :389-397 */
    uint64_t harm_ins = 0;
    for (uint32_t i = 0; i < t->n_inputs; i++) {
      const ok_utxo_entry* e = &entries[t->first_input + i];
      uint64_t p = plurality(e->script_len);
      if (e->amount == 0) return -1;
      harm_ins = satadd(harm_ins, C * p * p / e->amount); /* unchecked (wrapping) multiply as in the reference */
    }
    *mass = satsub(harm_outs, harm_ins);
    return 0;
  }
  uint64_t ins_plur = 0, sum_ins = 0;
  for (uint32_t i = 0; i < t->n_inputs; i++) { ins_plur += plurality(entries[t->first_input + i].script_len); sum_ins += entries[t->first_input + i].amount; }
  if (ins_plur == 0) return -1;
  uint64_t mean = sum_ins / ins_plur;
  if (mean == 0) return -1;
  *mass = satsub(harm_outs, satmul(ins_plur, C / mean));
  return 0;
}

/* ------------------------------------------------------------------ validate_populated_transaction_and_get_fee */
void ok_validate_populated(const ok_batch* b, const ok_utxo_entry* entries, size_t tx, uint64_t pov, int flags, const ok_params* p, ok_tx_result* out) {
  const ok_tx* t = &b->txs[tx];
  memset(out, 0, sizeof *out);
  /* coinbase maturity :75-91 */
  for (uint32_t i = 0; i < t->n_inputs; i++) {
    const ok_utxo_entry* e = &entries[t->first_input + i];
    if (e->is_coinbase && e->block_daa_score + p->coinbase_maturity > pov) { out->status = OK_TX_IMMATURE_COINBASE; out->fail_input = i; return; }
  }
  /* input amounts :93-108 */
  uint64_t total_in = 0;
  for (uint32_t i = 0; i < t->n_inputs; i++) {
    if (ckadd(total_in, entries[t->first_input + i].amount, &total_in)) { out->status = OK_TX_INPUT_AMOUNT_OVERFLOW; return; }
    if (total_in > p->max_sompi) { out->status = OK_TX_INPUT_AMOUNT_TOO_HIGH; return; }
  }
  /* outputs :110-118 (wrapping sum as in release Rust) */
  uint64_t total_out = 0;
  for (uint32_t i = 0; i < t->n_outputs; i++) total_out += b->outputs[t->first_output + i].value;
  if (total_in < total_out) { out->status = OK_TX_SPEND_TOO_HIGH; return; }
  out->fee = total_in - total_out;
  if (flags != OK_FLAGS_SKIP_MASS_CHECK) { /* :120-128 */
    uint64_t m;
    if (ok_storage_mass(b, entries, tx, p->storage_mass_parameter, &m)) { out->status = OK_TX_MASS_INCOMPUTABLE; return; }
    if (m != t->mass) { out->status = OK_TX_WRONG_MASS; return; }
  }
  /* sequence lock :130-155 */
  for (uint32_t i = 0; i < t->n_inputs; i++) {
    const ok_input* in = &b->inputs[t->first_input + i];
    if (in->sequence & (1ULL << 63)) continue;
    int64_t rel = (int64_t)(in->sequence & 0xffffffffULL);
    int64_t lock = (int64_t)entries[t->first_input + i].block_daa_score + rel - 1;
    if (lock >= (int64_t)pov) { out->status = OK_TX_SEQUENCE_LOCK; return; }
  }
  if (flags == OK_FLAGS_SKIP_SCRIPT_CHECKS) return;
  /* check_scripts :162-200 (sequential order => first failing input) */
  for (uint32_t i = 0; i < t->n_inputs; i++) {
    int err = ok_check_script_std(b, entries, tx, i);
    if (err == OK_SCRIPT_OK) continue;
    out->fail_input = i;
    out->script_err = (uint8_t)err;
    if (err == OK_SCRIPT_NONSTANDARD) out->status = OK_TX_NEEDS_HOST_VM;
    else out->status = b->inputs[t->first_input + i].sigscript_len == 0 ? OK_TX_SIGNATURE_EMPTY : OK_TX_SIGNATURE_INVALID; /* map_script_err :198-200 */
    return;
  }
}

/* ------------------------------------------------------------------ UTXO collections */
typedef struct { uint8_t key[36]; uint8_t used; /* 0 empty 1 full 2 tombstone */ ok_utxo_entry e; uint8_t* script; } slot_t;
typedef struct { slot_t* s; size_t cap, n, tomb; } map_t;

static uint64_t key_hash(const uint8_t k[36]) { uint64_t h; memcpy(&h, k, 8); uint32_t idx; memcpy(&idx, k + 32, 4); return (h ^ ((uint64_t)idx * 0x9E3779B97F4A7C15ULL)) * 0xD6E8FEB86659FD93ULL; }
static void map_init(map_t* m, size_t cap) { m->cap = cap; m->n = m->tomb = 0; m->s = calloc(cap, sizeof(slot_t)); }
static void map_free(map_t* m) { for (size_t i = 0; i < m->cap; i++) if (m->s[i].used == 1) free(m->s[i].script); free(m->s); }
static slot_t* map_find(const map_t* m, const uint8_t k[36]) {
  size_t i = key_hash(k) & (m->cap - 1);
  for (;;) {
    slot_t* s = &m->s[i];
    if (s->used == 0) return NULL;
    if (s->used == 1 && memcmp(s->key, k, 36) == 0) return s;
    i = (i + 1) & (m->cap - 1);
  }
}
static void map_put(map_t* m, const uint8_t k[36], const ok_utxo_entry* e, const uint8_t* script);
static void map_grow(map_t* m) {
  map_t n;
  map_init(&n, m->cap * 2);
  for (size_t i = 0; i < m->cap; i++) if (m->s[i].used == 1) { map_put(&n, m->s[i].key, &m->s[i].e, m->s[i].script); }
  map_free(m);
  *m = n;
}
static void map_put(map_t* m, const uint8_t k[36], const ok_utxo_entry* e, const uint8_t* script) {
  if ((m->n + m->tomb + 1) * 2 > m->cap) map_grow(m);
  slot_t* f = map_find(m, k);
  if (!f) {
    size_t i = key_hash(k) & (m->cap - 1);
    while (m->s[i].used == 1) i = (i + 1) & (m->cap - 1);
    f = &m->s[i];
    if (f->used == 2) m->tomb--;
    memcpy(f->key, k, 36);
    f->used = 1;
    m->n++;
  } else {
    free(f->script);
  }
  f->e = *e;
  f->script = malloc(e->script_len ? e->script_len : 1);
  memcpy(f->script, script, e->script_len);
}
static void map_del(map_t* m, const uint8_t k[36]) {
  slot_t* f = map_find(m, k);
  if (!f) return;
  free(f->script); f->script = NULL; f->used = 2; m->n--; m->tomb++;
}

struct ok_state { map_t base, add, rem; };
ok_state* ok_state_new(void) { ok_state* s = malloc(sizeof *s); map_init(&s->base, 1024); map_init(&s->add, 256); map_init(&s->rem, 256); return s; }
void ok_state_free(ok_state* s) { map_free(&s->base); map_free(&s->add); map_free(&s->rem); free(s); }

/* ComposedUtxoView::get (utxo_view.rs:22-35) */
static const slot_t* state_get(const ok_state* s, const uint8_t k[36]) {
  const slot_t* f = map_find(&s->add, k);
  if (f) return f;
  if (map_find(&s->rem, k)) return NULL;
  return map_find(&s->base, k);
}
int ok_state_get(const ok_state* s, const uint8_t key36[36], ok_utxo_entry* e, uint8_t* script, size_t cap) {
  const slot_t* f = state_get(s, key36);
  if (!f) return 0;
  if (e) *e = f->e;
  if (script) memcpy(script, f->script, f->e.script_len < cap ? f->e.script_len : cap);
  return 1;
}
uint64_t ok_state_count(const ok_state* s) {
  uint64_t n = s->add.n;
  for (size_t i = 0; i < s->base.cap; i++)
    if (s->base.s[i].used == 1 && !map_find(&s->add, s->base.s[i].key) && !map_find(&s->rem, s->base.s[i].key)) n++;
  return n;
}

static void make_key(uint8_t k[36], const uint8_t txid[32], uint32_t index) { memcpy(k, txid, 32); for (int i = 0; i < 4; i++) k[32 + i] = (uint8_t)(index >> (8 * i)); }

typedef struct { ok_state* s; const ok_batch* b; uint64_t pov; int flags; const ok_params* p; ok_tx_result* res; size_t lo, hi; } vjob;
static void* vworker(void* a) {
  vjob* j = (vjob*)a;
  const ok_batch* b = j->b;
  for (size_t ti = j->lo; ti < j->hi; ti++) {
    const ok_tx* t = &b->txs[ti];
    ok_tx_result* r = &j->res[ti];
    memset(r, 0, sizeof *r);
    if (tx_is_coinbase(t)) { r->status = OK_TX_SKIPPED_COINBASE; continue; }
    /* populate (utxo_validation.rs:319-327): a private single-tx batch whose entry scripts live in a scratch arena */
    size_t need = 0;
    int missing = 0;
    const slot_t* found[1024];
    ok_utxo_entry* ents = calloc(b->n_inputs ? b->n_inputs : 1, sizeof(ok_utxo_entry)); /* indexed absolutely like b->inputs */
    const slot_t** fs = t->n_inputs <= 1024 ? found : malloc(sizeof(slot_t*) * t->n_inputs);
    for (uint32_t i = 0; i < t->n_inputs && !missing; i++) {
      const ok_input* in = &b->inputs[t->first_input + i];
      uint8_t k[36];
      make_key(k, in->prev_txid, in->prev_index);
      fs[i] = state_get(j->s, k);
      if (!fs[i]) missing = 1; else need += fs[i]->e.script_len;
    }
    if (missing) { r->status = OK_TX_MISSING_OUTPOINTS; }
    else {
      /* entry scripts must be addressable through b->bytes offsets: build a shadow arena = b->bytes ++ scripts */
      uint8_t* arena = malloc(b->n_bytes + need + 1);
      memcpy(arena, b->bytes, b->n_bytes);
      size_t off = b->n_bytes;
      for (uint32_t i = 0; i < t->n_inputs; i++) {
        ents[t->first_input + i] = fs[i]->e;
        ents[t->first_input + i].script_off = (uint32_t)off;
        memcpy(arena + off, fs[i]->script, fs[i]->e.script_len);
        off += fs[i]->e.script_len;
      }
      ok_batch sb = *b;
      sb.bytes = arena; sb.n_bytes = off;
      ok_validate_populated(&sb, ents, ti, j->pov, j->flags, j->p, r);
      free(arena);
    }
    if (fs != found) free(fs);
    free(ents);
  }
  return NULL;
}
void ok_state_validate(ok_state* s, const ok_batch* b, uint64_t pov, int flags, const ok_params* p, ok_tx_result* results, int nthreads) {
  ok_secp_init();
  size_t n = b->n_txs;
  if (nthreads < 1) nthreads = 1;
  if ((size_t)nthreads > n) nthreads = n ? (int)n : 1;
  pthread_t* th = malloc(sizeof(pthread_t) * nthreads);
  vjob* jobs = malloc(sizeof(vjob) * nthreads);
  for (int t = 0; t < nthreads; t++) {
    jobs[t] = (vjob){s, b, pov, flags, p, results, n * t / nthreads, n * (t + 1) / nthreads};
    if (t + 1 < nthreads) pthread_create(&th[t], NULL, vworker, &jobs[t]);
  }
  vworker(&jobs[nthreads - 1]);
  for (int t = 0; t + 1 < nthreads; t++) pthread_join(th[t], NULL);
  free(th); free(jobs);
}

/* UtxoDiff::add_transaction (utxo_diff.rs:233-269) */
static int accept_range(ok_state* s, const ok_batch* b, size_t t0, size_t t1, const uint8_t* accept, uint64_t pov);
int ok_state_accept(ok_state* s, const ok_batch* b, const uint8_t* accept, uint64_t pov) { return accept_range(s, b, 0, b->n_txs, accept, pov); }
static int accept_range(ok_state* s, const ok_batch* b, size_t t0, size_t t1, const uint8_t* accept, uint64_t pov) {
  for (size_t ti = t0; ti < t1; ti++) {
    if (!accept[ti]) continue;
    const ok_tx* t = &b->txs[ti];
    for (uint32_t i = 0; i < t->n_inputs; i++) { /* remove_entry */
      const ok_input* in = &b->inputs[t->first_input + i];
      uint8_t k[36];
      make_key(k, in->prev_txid, in->prev_index);
      const slot_t* cur = state_get(s, k);
      if (!cur) return -1;
      slot_t* a = map_find(&s->add, k);
      if (a && a->e.block_daa_score == cur->e.block_daa_score) map_del(&s->add, k);
      else if (!map_find(&s->rem, k)) { ok_utxo_entry e = cur->e; uint8_t* sc = malloc(e.script_len + 1); memcpy(sc, cur->script, e.script_len); map_put(&s->rem, k, &e, sc); free(sc); }
      else return -1; /* DoubleRemoveCall */
    }
    uint8_t id[32];
    ok_tx_id(b, ti, id);
    int cb = tx_is_coinbase(t);
    for (uint32_t i = 0; i < t->n_outputs; i++) { /* add_entry */
      const ok_output* o = &b->outputs[t->first_output + i];
      uint8_t k[36];
      make_key(k, id, i);
      ok_utxo_entry e;
      memset(&e, 0, sizeof e);
      e.amount = o->value; e.block_daa_score = pov; e.script_len = o->script_len; e.spk_version = o->spk_version; e.is_coinbase = (uint8_t)cb;
      slot_t* r = map_find(&s->rem, k);
      if (r && r->e.block_daa_score == pov) map_del(&s->rem, k);
      else if (!map_find(&s->add, k)) map_put(&s->add, k, &e, b->bytes + o->script_off);
      else return -1; /* DoubleAddCall */
    }
  }
  return 0;
}

/* --------------------------------------------------------------------------------------------------------------
 * calculate_utxo_state as ONE call over a window of blocks (utxo_validation.rs:110-173): for every block in order
 *   validate_transactions_in_parallel (tx 0 = coinbase, skipped by position :273) on a pool of worker threads,
 *   UtxoDiff::add_transaction for the accepted ones (+ the coinbase when block_flags bit 0 is set; NULL = always), commit.
 * Block b = transactions [block_first_tx[b], block_first_tx[b+1]).  block_flags bit 1 = SkipScriptChecks, bit 2 = verify only.
 * This is the CPU baseline of the DAG-replay benchmark and the checker of kgv_replay_window.
 * Workers split a block's transactions into static contiguous chunks (rayon par_iter over a Vec); the pool lives for the
 * whole call.  Entries are populated by the calling thread before the fan-out (HashMap probes are noise next to the
 * signature checks) into ONE shadow arena = window bytes ++ spent scripts.
 * -------------------------------------------------------------------------------------------------------------- */
typedef struct {
  pthread_mutex_t mu; pthread_cond_t go, done;
  int nthreads, generation, remaining, quit;
  /* current job */
  const ok_batch* sb; const ok_utxo_entry* ents; const uint8_t* missing; uint64_t pov; int flags; const ok_params* p; ok_tx_result* res; size_t t0, t1;
} rpool;
typedef struct { rpool* pool; int id; } rworker_arg;
static void replay_chunk(rpool* q, int id) {
  size_t n = q->t1 - q->t0, lo = q->t0 + n * (size_t)id / (size_t)q->nthreads, hi = q->t0 + n * ((size_t)id + 1) / (size_t)q->nthreads;
  for (size_t ti = lo; ti < hi; ti++) {
    ok_tx_result* r = &q->res[ti];
    memset(r, 0, sizeof *r);
    if (ti == q->t0 || tx_is_coinbase(&q->sb->txs[ti])) { r->status = OK_TX_SKIPPED_COINBASE; continue; }
    if (q->missing[ti]) { r->status = OK_TX_MISSING_OUTPOINTS; continue; }
    ok_validate_populated(q->sb, q->ents, ti, q->pov, q->flags, q->p, r);
  }
}
static void* rworker(void* a) {
  rworker_arg* w = (rworker_arg*)a;
  rpool* q = w->pool;
  int seen = 0;
  for (;;) {
    pthread_mutex_lock(&q->mu);
    while (q->generation == seen && !q->quit) pthread_cond_wait(&q->go, &q->mu);
    if (q->quit) { pthread_mutex_unlock(&q->mu); return NULL; }
    seen = q->generation;
    pthread_mutex_unlock(&q->mu);
    replay_chunk(q, w->id);
    pthread_mutex_lock(&q->mu);
    if (--q->remaining == 0) pthread_cond_signal(&q->done);
    pthread_mutex_unlock(&q->mu);
  }
}
static int accept_range(ok_state* s, const ok_batch* b, size_t t0, size_t t1, const uint8_t* accept, uint64_t pov);

int ok_state_replay(ok_state* s, const ok_batch* b, const uint32_t* block_first_tx, const uint64_t* block_pov, const uint32_t* block_flags, size_t n_blocks,
                    const ok_params* p, ok_tx_result* results, uint8_t* accept_out, int nthreads) {
  ok_secp_init();
  if (nthreads < 1) nthreads = 1;
  rpool q;
  memset(&q, 0, sizeof q);
  pthread_mutex_init(&q.mu, NULL); pthread_cond_init(&q.go, NULL); pthread_cond_init(&q.done, NULL);
  q.nthreads = nthreads; q.p = p; q.res = results;
  pthread_t* th = malloc(sizeof(pthread_t) * (size_t)nthreads);
  rworker_arg* wa = malloc(sizeof(rworker_arg) * (size_t)nthreads);
  for (int t = 1; t < nthreads; t++) { wa[t].pool = &q; wa[t].id = t; pthread_create(&th[t], NULL, rworker, &wa[t]); }
  size_t cap = b->n_bytes + 64 * b->n_inputs + 4096, off = b->n_bytes;
  uint8_t* shadow = malloc(cap);
  memcpy(shadow, b->bytes, b->n_bytes);
  ok_utxo_entry* ents = calloc(b->n_inputs ? b->n_inputs : 1, sizeof(ok_utxo_entry));
  uint8_t* missing = calloc(b->n_txs ? b->n_txs : 1, 1);
  uint8_t* acc = accept_out ? accept_out : calloc(b->n_txs ? b->n_txs : 1, 1);
  int rc = 0;
  for (size_t bi = 0; bi < n_blocks && rc == 0; bi++) {
    size_t t0 = block_first_tx[bi], t1 = block_first_tx[bi + 1];
    uint32_t bf = block_flags ? block_flags[bi] : 1u;
    if (t1 <= t0) continue;
    for (size_t ti = t0 + 1; ti < t1; ti++) { /* populate: utxo_validation.rs:319-327 */
      const ok_tx* t = &b->txs[ti];
      missing[ti] = 0;
      for (uint32_t i = 0; i < t->n_inputs; i++) {
        const ok_input* in = &b->inputs[t->first_input + i];
        uint8_t k[36];
        make_key(k, in->prev_txid, in->prev_index);
        const slot_t* f = state_get(s, k);
        if (!f) { missing[ti] = 1; break; }
        if (off + f->e.script_len > cap) { cap = 2 * cap + f->e.script_len; shadow = realloc(shadow, cap); }
        ents[t->first_input + i] = f->e;
        ents[t->first_input + i].script_off = (uint32_t)off;
        memcpy(shadow + off, f->script, f->e.script_len);
        off += f->e.script_len;
      }
    }
    ok_batch sb = *b;
    sb.bytes = shadow; sb.n_bytes = off;
    pthread_mutex_lock(&q.mu);
    q.sb = &sb; q.ents = ents; q.missing = missing; q.pov = block_pov[bi]; q.flags = (bf & 2u) ? OK_FLAGS_SKIP_SCRIPT_CHECKS : OK_FLAGS_FULL; q.t0 = t0; q.t1 = t1;
    q.remaining = nthreads - 1; q.generation++;
    pthread_cond_broadcast(&q.go);
    pthread_mutex_unlock(&q.mu);
    replay_chunk(&q, 0);
    pthread_mutex_lock(&q.mu);
    while (q.remaining) pthread_cond_wait(&q.done, &q.mu);
    pthread_mutex_unlock(&q.mu);
    for (size_t ti = t0; ti < t1; ti++) acc[ti] = (uint8_t)(!(bf & 4u) && (ti == t0 ? (bf & 1u) != 0 : results[ti].status == OK_TX_OK));
    if (accept_range(s, b, t0, t1, acc, block_pov[bi]) != 0) rc = -1;
    ok_state_commit(s);
    off = b->n_bytes; /* the spent scripts of this block are no longer needed */
  }
  pthread_mutex_lock(&q.mu); q.quit = 1; pthread_cond_broadcast(&q.go); pthread_mutex_unlock(&q.mu);
  for (int t = 1; t < nthreads; t++) pthread_join(th[t], NULL);
  free(th); free(wa); free(shadow); free(ents); free(missing);
  if (!accept_out) free(acc);
  pthread_mutex_destroy(&q.mu); pthread_cond_destroy(&q.go); pthread_cond_destroy(&q.done);
  return rc;
}

void ok_state_commit(ok_state* s) { /* write_diff_batch: delete removed, then put added */
  for (size_t i = 0; i < s->rem.cap; i++) if (s->rem.s[i].used == 1) map_del(&s->base, s->rem.s[i].key);
  for (size_t i = 0; i < s->add.cap; i++) if (s->add.s[i].used == 1) map_put(&s->base, s->add.s[i].key, &s->add.s[i].e, s->add.s[i].script);
  map_free(&s->add); map_free(&s->rem);
  map_init(&s->add, 256); map_init(&s->rem, 256);
}

static void digest_add(uint64_t acc[4], const slot_t* f) {
  ok_blake2b_ctx h;
  ok_blake2b_init(&h, 32, "MuHashElement", 13);
  ok_blake2b_update(&h, f->key, 36);
  uint8_t tmp[8];
  for (int i = 0; i < 8; i++) tmp[i] = (uint8_t)(f->e.block_daa_score >> (8 * i));
  ok_blake2b_update(&h, tmp, 8);
  for (int i = 0; i < 8; i++) tmp[i] = (uint8_t)(f->e.amount >> (8 * i));
  ok_blake2b_update(&h, tmp, 8);
  tmp[0] = f->e.is_coinbase ? 1 : 0;
  ok_blake2b_update(&h, tmp, 1);
  tmp[0] = (uint8_t)f->e.spk_version; tmp[1] = (uint8_t)(f->e.spk_version >> 8);
  ok_blake2b_update(&h, tmp, 2);
  uint64_t l = f->e.script_len;
  for (int i = 0; i < 8; i++) tmp[i] = (uint8_t)(l >> (8 * i));
  ok_blake2b_update(&h, tmp, 8);
  ok_blake2b_update(&h, f->script, f->e.script_len);
  uint8_t d[32];
  ok_blake2b_final(&h, d);
  unsigned __int128 c = 0;
  for (int i = 0; i < 4; i++) { uint64_t w; memcpy(&w, d + 8 * i, 8); c += (unsigned __int128)acc[i] + w; acc[i] = (uint64_t)c; c >>= 64; }
}
void ok_state_digest(const ok_state* s, uint8_t out[32]) {
  uint64_t acc[4] = {0, 0, 0, 0};
  for (size_t i = 0; i < s->add.cap; i++) if (s->add.s[i].used == 1) digest_add(acc, &s->add.s[i]);
  for (size_t i = 0; i < s->base.cap; i++)
    if (s->base.s[i].used == 1 && !map_find(&s->add, s->base.s[i].key) && !map_find(&s->rem, s->base.s[i].key)) digest_add(acc, &s->base.s[i]);
  memcpy(out, acc, 32);
}
