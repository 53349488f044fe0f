/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the shipped product path.
 *
 * CPU restatement (plain C11) of the rusty-kaspa transaction-validation hot path
 * (SURVEY.md §8a), used as the parity checker for the CUDA kernels and as the
 * "port" CPU baseline.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it.  Every function cites the
 * reference file:line it follows.
 *
 * Parity pinning: hashes, tx-id/tx-hash, sighash, the mainnet Schnorr P2PK and
 * 2-of-4 P2SH multisig KATs and the 224-input simpa fixture are pinned by the
 * reference's own vectors (tests/golden/ *.json, tests/test_oracle_golden.py);
 * MuHash by every known answer of crypto/muhash/src/lib.rs (tests/golden/muhash.json,
 * tests/test_oracle_muhash.py); tx hash + merkle root by the hashMerkleRoot of all 266
 * headers of the simpa DAG fixture.
 * ECDSA verdicts and the Schnorr edge encodings (r>=p, s>=n, off-curve pk) are
 * NOT covered by any stored vector in the reference: for those "parity
 * unpinned" — they are cross-checked against an independent big-int restatement
 * (oracle/pyref.py) and OpenSSL (`cryptography`) only.  Likewise "parity unpinned":
 * the ERROR cases of the block-body set checks (ok_block_set_checks): the reference's
 * test blocks for them are Rust literals; the passing case is pinned by the fixture.
 */
#ifndef OK_ORACLE_H
#define OK_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- hashes (ok_hash.c) ---- */
typedef struct { uint32_t st[8]; uint64_t len; uint8_t buf[64]; } ok_sha256_ctx;
void ok_sha256_init(ok_sha256_ctx* c);
void ok_sha256_update(ok_sha256_ctx* c, const void* data, size_t n);
void ok_sha256_final(ok_sha256_ctx* c, uint8_t out[32]);
void ok_sha256(const void* data, size_t n, uint8_t out[32]);
void ok_sha256_domain(const char* domain, const void* data, size_t n, uint8_t out[32]);

typedef struct { uint64_t h[8]; uint64_t t[2]; uint8_t buf[128]; size_t fill; size_t outlen; } ok_blake2b_ctx;
void ok_blake2b_init(ok_blake2b_ctx* c, size_t outlen, const void* key, size_t keylen);
void ok_blake2b_update(ok_blake2b_ctx* c, const void* data, size_t n);
void ok_blake2b_final(ok_blake2b_ctx* c, uint8_t* out);
void ok_blake2b_keyed(const char* domain, const void* data, size_t n, uint8_t out[32]);
void ok_blake2b_256(const void* data, size_t n, uint8_t out[32]);

/* ---- secp256k1 (ok_secp256k1.c) ---- */
/* Per-item verdicts: the reference distinguishes a parse error
 * (script aborts with InvalidSignature) from a well-formed but wrong signature. */
enum { OK_SIG_INVALID = 0, OK_SIG_VALID = 1, OK_SIG_PK_PARSE_ERR = 2, OK_SIG_SIG_PARSE_ERR = 3 };

void ok_secp_init(void); /* builds the generator table once; idempotent, thread-safe after first call */
int ok_schnorr_verify(const uint8_t pk32[32], const uint8_t msg32[32], const uint8_t sig64[64]);
int ok_ecdsa_verify(const uint8_t pk33[33], const uint8_t msg32[32], const uint8_t sig64[64]);
/* batch forms: SoA, nthreads >= 1 worker threads with contiguous static chunks
 * (mirrors rayon par_iter over a slice, consensus/src/pipeline/virtual_processor/utxo_validation.rs:269-277) */
void ok_schnorr_verify_batch(const uint8_t* pk32, const uint8_t* msg32, const uint8_t* sig64, size_t n, uint8_t* status, int nthreads);
void ok_ecdsa_verify_batch(const uint8_t* pk33, const uint8_t* msg32, const uint8_t* sig64, size_t n, uint8_t* status, int nthreads);

/* "cpu_fast" (ok_secp_fast.c): the speed-oriented second port (GLV, wNAF, effective-affine tables, dedicated squaring) that the
 * CPU baselines time; verdicts identical to the plain checker's above (tests/test_oracle_secp.py). */
int ok_schnorr_verify_fast(const uint8_t pk32[32], const uint8_t msg32[32], const uint8_t sig64[64]);
int ok_ecdsa_verify_fast(const uint8_t pk33[33], const uint8_t msg32[32], const uint8_t sig64[64]);
void ok_schnorr_verify_batch_fast(const uint8_t* pk32, const uint8_t* msg32, const uint8_t* sig64, size_t n, uint8_t* status, int nthreads);
void ok_ecdsa_verify_batch_fast(const uint8_t* pk33, const uint8_t* msg32, const uint8_t* sig64, size_t n, uint8_t* status, int nthreads);
/* route the signature checks of ok_validate.c (ok_state_validate / ok_state_replay) through the fast port: on for the CPU baselines,
 * off (default) when the oracle acts as the checker */
void ok_use_fast_verify(int on);

/* test-vector generation helpers (BIP-340 signing with aux=0^32; ECDSA with a
 * SHA-256 derived nonce and low-S normalisation).  Return 1 on success. */
int ok_schnorr_pubkey(const uint8_t seckey32[32], uint8_t pk32[32]);
int ok_schnorr_sign(const uint8_t seckey32[32], const uint8_t msg32[32], uint8_t sig64[64]);
int ok_ecdsa_pubkey(const uint8_t seckey32[32], uint8_t pk33[33]);
int ok_ecdsa_sign(const uint8_t seckey32[32], const uint8_t msg32[32], uint8_t sig64[64]);
/* low-level probes used by the cross-checks against oracle/pyref.py */
int ok_ec_mul_xy(const uint8_t scalar32[32], const uint8_t px32[32], const uint8_t py32[32], uint8_t outx[32], uint8_t outy[32]);
void ok_fe_mul_bytes(const uint8_t a[32], const uint8_t b[32], uint8_t out[32]);
void ok_sc_mul_bytes(const uint8_t a[32], const uint8_t b[32], uint8_t out[32]);

/* ---- transaction model (ok_tx.c) ----
 * Flat SoA batch: the same memory layout the product ABI uses (include/kgv.h kgv_tx / kgv_input /
 * kgv_output / kgv_utxo_entry), restated here so that tests can hand ONE set of numpy arrays to both
 * sides.  Mirrors consensus/core/src/tx.rs:49-57,72-77,93-101,121-125,165-185. */
typedef struct {
  uint32_t first_input, n_inputs, first_output, n_outputs;
  uint64_t lock_time, gas, mass;
  uint32_t payload_off, payload_len;
  uint16_t version;
  uint8_t subnetwork_id[20];
  uint8_t flags; /* unused by the oracle: coinbase-ness is derived from subnetwork_id */
  uint8_t pad_;
} ok_tx; /* 72 bytes */
typedef struct {
  uint8_t prev_txid[32];
  uint32_t prev_index;
  uint32_t sigscript_off, sigscript_len;
  uint8_t sig_op_count;
  uint8_t pad_[3];
  uint64_t sequence;
} ok_input; /* 56 bytes */
typedef struct {
  uint64_t value;
  uint32_t script_off, script_len;
  uint16_t spk_version;
  uint8_t pad_[6];
} ok_output; /* 24 bytes */
typedef struct {
  uint64_t amount;
  uint64_t block_daa_score;
  uint32_t script_off, script_len;
  uint16_t spk_version;
  uint8_t is_coinbase;
  uint8_t pad_[5];
} ok_utxo_entry; /* 32 bytes */
typedef struct {
  const ok_tx* txs; size_t n_txs;
  const ok_input* inputs; size_t n_inputs;
  const ok_output* outputs; size_t n_outputs;
  const uint8_t* bytes; size_t n_bytes;
} ok_batch;

/* consensus/core/src/hashing/tx.rs:16-107 */
void ok_tx_id(const ok_batch* b, size_t tx, uint8_t out[32]);
void ok_tx_hash(const ok_batch* b, size_t tx, uint8_t out[32]);
void ok_tx_ids(const ok_batch* b, uint8_t* out32, int nthreads);
void ok_tx_hashes(const ok_batch* b, uint8_t* out32, int nthreads);
/* crypto/merkle/src/lib.rs:3-30 (calc_merkle_root over n 32-byte hashes; n == 0 -> ZERO_HASH) */
void ok_merkle_root(const uint8_t* hashes32, size_t n, uint8_t out[32]);
/* body_validation_in_isolation.rs:95-131 for the block [t0, t1) of the batch (see ok_tx.c) */
int ok_block_set_checks(const ok_batch* b, uint32_t t0, uint32_t t1, uint32_t* index);
/* consensus/core/src/hashing/sighash.rs:140-277; entries[] is indexed like b->inputs (one populated
 * UTXO entry per input, scripts in b->bytes); input_index is relative to the tx. */
void ok_sighash(const ok_batch* b, const ok_utxo_entry* entries, size_t tx, uint32_t input_index, uint8_t hash_type, int ecdsa, uint8_t out[32]);

/* ---- UTXO state, transaction validation in UTXO context (ok_validate.c) ----
 * Status / error numbering is shared with include/kgv.h (KGV_TX_*, KGV_SCRIPT_*). */
enum { OK_TX_OK = 0, OK_TX_MISSING_OUTPOINTS = 1, OK_TX_IMMATURE_COINBASE = 2, OK_TX_INPUT_AMOUNT_OVERFLOW = 3,
       OK_TX_INPUT_AMOUNT_TOO_HIGH = 4, OK_TX_SPEND_TOO_HIGH = 5, OK_TX_MASS_INCOMPUTABLE = 6, OK_TX_WRONG_MASS = 7,
       OK_TX_SEQUENCE_LOCK = 8, OK_TX_SIGNATURE_INVALID = 9, OK_TX_SIGNATURE_EMPTY = 10, OK_TX_NEEDS_HOST_VM = 11, OK_TX_SKIPPED_COINBASE = 12 };
enum { OK_SCRIPT_OK = 0, OK_SCRIPT_EVAL_FALSE = 1, OK_SCRIPT_NULL_FAIL = 2, OK_SCRIPT_INVALID_SIGNATURE = 3, OK_SCRIPT_SIG_LENGTH = 4,
       OK_SCRIPT_PUBKEY_FORMAT = 5, OK_SCRIPT_INVALID_SIGHASH_TYPE = 6, OK_SCRIPT_EXCEEDED_SIGOP_LIMIT = 7, OK_SCRIPT_NONSTANDARD = 255 };
enum { OK_FLAGS_FULL = 0, OK_FLAGS_SKIP_SCRIPT_CHECKS = 1, OK_FLAGS_SKIP_MASS_CHECK = 2 };
typedef struct { uint64_t coinbase_maturity, storage_mass_parameter, max_sompi; } ok_params;
typedef struct { uint64_t fee; uint32_t fail_input; uint8_t status, script_err, pad_[2]; } ok_tx_result; /* 16 bytes */

/* Script check of ONE input for the standard script classes (P2PK Schnorr, P2PK ECDSA, P2SH m-of-n
 * multisig), following TxScriptEngine::execute for exactly these shapes (crypto/txscript/src/lib.rs:399-643,
 * opcodes/mod.rs:746-808, script_class.rs:58-82).  Anything else returns OK_SCRIPT_NONSTANDARD. */
int ok_check_script_std(const ok_batch* b, const ok_utxo_entry* entries, size_t tx, uint32_t input_index);

/* storage mass, consensus/core/src/mass/mod.rs:338-410. returns 0 and *mass on success, -1 if incomputable */
int ok_storage_mass(const ok_batch* b, const ok_utxo_entry* entries, size_t tx, uint64_t storm_param, uint64_t* mass);

/* validate_populated_transaction_and_get_fee (tx_validation_in_utxo_context.rs:34-61) with entries given. */
void ok_validate_populated(const ok_batch* b, const ok_utxo_entry* entries, size_t tx, uint64_t pov_daa_score, int flags, const ok_params* p, ok_tx_result* out);

/* UTXO state = base set composed with one diff layer (utxo_view.rs:22-35, utxo_diff.rs:15-19,233-269,
 * model/stores/utxo_set.rs:107-112). */
typedef struct ok_state ok_state;
ok_state* ok_state_new(void);
void ok_state_free(ok_state* s);
/* validate_transactions_in_parallel (utxo_validation.rs:262-338): every non-coinbase tx of the batch against
 * the composed view; results[i] for tx i (coinbase: OK_TX_SKIPPED_COINBASE). */
void ok_state_validate(ok_state* s, const ok_batch* b, uint64_t pov_daa_score, int flags, const ok_params* p, ok_tx_result* results, int nthreads);
/* mergeset_diff.add_transaction for every tx with accept[i] != 0 (utxo_diff.rs:233-247). returns 0, or -1 on a UtxoAlgebraError */
int ok_state_accept(ok_state* s, const ok_batch* b, const uint8_t* accept, uint64_t pov_daa_score);
/* calculate_utxo_state over a window of blocks in one call (utxo_validation.rs:110-173): per block validate in parallel on a
 * persistent pool of nthreads workers (tx 0 = coinbase, skipped by position), accept, commit.  block_flags (may be NULL = 1):
 * bit 0 accept the coinbase, bit 1 SkipScriptChecks, bit 2 validate only.  accept_out may be NULL.  Returns 0, or -1 on a
 * UtxoAlgebraError. */
int ok_state_replay(ok_state* s, const ok_batch* b, const uint32_t* block_first_tx, const uint64_t* block_pov, const uint32_t* block_flags, size_t n_blocks,
                    const ok_params* p, ok_tx_result* results, uint8_t* accept_out, int nthreads);
/* write_diff_batch: fold the diff into the base (utxo_set.rs:107-112) */
void ok_state_commit(ok_state* s);
/* composed get: returns 1 if found (entry header + script bytes copied, script_cap bytes max) */
int ok_state_get(const ok_state* s, const uint8_t key36[36], ok_utxo_entry* e, uint8_t* script, size_t script_cap);
uint64_t ok_state_count(const ok_state* s);
/* order-independent digest of the composed set: sum mod 2^256 of the MuHashElement hashes (consensus/core/src/muhash.rs:47-59) */
void ok_state_digest(const ok_state* s, uint8_t out[32]);

/* ---- MuHash (ok_muhash.c): crypto/muhash/src/lib.rs, u3072.rs, consensus/core/src/muhash.rs ---- */
typedef struct { uint64_t num[48], den[48]; } ok_muhash; /* canonical residues modulo 2^3072 - 1103717 */
void ok_muhash_init(ok_muhash* m);
void ok_muhash_expand(const uint8_t hash32[32], uint8_t out384[384]);
void ok_muhash_add_element(ok_muhash* m, const void* data, size_t n);
void ok_muhash_remove_element(ok_muhash* m, const void* data, size_t n);
void ok_muhash_combine(ok_muhash* m, const ok_muhash* o);
void ok_muhash_serialize(ok_muhash* m, uint8_t out384[384]);
void ok_muhash_finalize(ok_muhash* m, uint8_t out32[32]);
int ok_muhash_deserialize(ok_muhash* m, const uint8_t in384[384]);
void ok_muhash_raw(const ok_muhash* m, uint8_t num384[384], uint8_t den384[384]);
void ok_muhash_add_utxo(ok_muhash* m, const uint8_t key36[36], const ok_utxo_entry* e, const uint8_t* bytes);
void ok_muhash_add_transaction(ok_muhash* m, const ok_batch* b, const ok_utxo_entry* entries, size_t tx, uint64_t block_daa_score);
void ok_muhash_accepted(ok_muhash* m, const ok_batch* b, const ok_utxo_entry* entries, const uint8_t* accept, uint64_t pov_daa_score);

#ifdef __cplusplus
}
#endif
#endif
