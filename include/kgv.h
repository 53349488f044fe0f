/*
 * kgv.h — C ABI of the B200 transaction-validation library (libkgv.so).
 *
 * "kgv" = Kaspa GPU Validator.  This is the drop-in boundary for the hot path named by
 * BASELINE.json: batched secp256k1 Schnorr/ECDSA verification, sighash / tx-id hashing, the
 * UTXO table and the fused per-transaction validation that rusty-kaspa fans out over rayon in
 *   consensus/src/pipeline/virtual_processor/utxo_validation.rs:262-309
 *   consensus/src/processes/transaction_validator/tx_validation_in_utxo_context.rs:34-61,157-196
 *   crypto/txscript/src/lib.rs:574-643
 * The reference has no FFI layer for this path: these entry points are what a
 * Rust shim (`extern "C"` block, shown in INTEGRATION.md) binds to stand behind the reference's
 * own TransactionValidator / SigCache / UtxoView surface.
 *
 * Conventions
 *   - plain pointers and sizes only; caller owns every buffer; the library never keeps a host
 *     pointer past the call.
 *   - every data pointer may be a HOST pointer (pageable or pinned) or a DEVICE pointer on the
 *     context's device; the library detects which (cudaPointerGetAttributes).  All data
 *     pointers of one call must be of the same kind.  With host pointers the call copies in,
 *     computes, copies out and returns after the results are in the caller's memory.  With
 *     device pointers the work is enqueued on the context's stream and the call returns
 *     without synchronising (use kgv_synchronize or your own stream sync).
 *   - return value: 0 = ok, negative = argument / CUDA / NCCL failure (kgv_last_error explains).
 *     An invalid signature is NEVER an error return: verdicts are per-item status bytes.
 *   - there is no CPU fallback: without a usable CUDA device kgv_create fails.
 *   - results are deterministic and independent of batch split or GPU count.
 *   - a context serialises its calls with an internal mutex; use one context per thread for
 *     concurrency.
 */
#ifndef KGV_H
#define KGV_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KGV_OK 0
#define KGV_ERR_ARG (-1)
#define KGV_ERR_CUDA (-2)
#define KGV_ERR_NOMEM (-3)
#define KGV_ERR_NCCL (-4)
#define KGV_ERR_LIMIT (-5) /* an internal iteration cap was hit (kgv_check_scripts_host); kgv_last_error says which */

/* Per-signature verdicts.  The reference distinguishes these cases
 * (crypto/txscript/src/lib.rs:582-583, 593, 618-619, 628):
 *   a malformed key / overflowing ECDSA r|s aborts the script with InvalidSignature,
 *   a well-formed but wrong signature is Ok(false) and (in multisig) the loop moves on. */
#define KGV_SIG_INVALID 0       /* sig.verify(..) -> Err  => Ok(false)                       */
#define KGV_SIG_VALID 1         /* sig.verify(..) -> Ok   => Ok(true)                        */
#define KGV_SIG_PK_PARSE_ERR 2  /* XOnlyPublicKey::from_slice / PublicKey::from_slice failed */
#define KGV_SIG_SIG_PARSE_ERR 3 /* ecdsa::Signature::from_compact failed (r or s >= n)       */

typedef struct kgv_ctx kgv_ctx;

/* One context per device.  Builds the generator window tables (8 MiB, L2 resident) on the GPU. */
int kgv_create(int device, uint32_t flags, kgv_ctx** out);
void kgv_destroy(kgv_ctx* ctx);
/* Run all subsequent work of this context on the given cudaStream_t (NULL = CUDA's default
 * stream, as for any cudaStream_t).  Lets a caller bracket kernels with its own events (bench.py
 * passes torch's stream).  kgv_reset_stream goes back to the context's private stream. */
int kgv_set_stream(kgv_ctx* ctx, void* cuda_stream);
int kgv_reset_stream(kgv_ctx* ctx);
int kgv_synchronize(kgv_ctx* ctx);
const char* kgv_last_error(const kgv_ctx* ctx);
/* Number of kernel launches this context has issued so far (bench.py's gpu_launches). */
uint64_t kgv_launch_count(const kgv_ctx* ctx);

/* Batched BIP-340 Schnorr verification: status[i] = verdict of (pk32[i], msg32[i], sig64[i]).
 * Replaces the per-signature FFI call `sig.verify(&msg,&pk)` of
 * crypto/txscript/src/lib.rs:593 together with the parses at :582-583.
 * SoA layout: pk32 = n*32 bytes (x-only key), msg32 = n*32 bytes, sig64 = n*64 bytes (r||s). */
int kgv_schnorr_verify(kgv_ctx* ctx, const uint8_t* pk32, const uint8_t* msg32, const uint8_t* sig64, size_t n, uint8_t* status);

/* Batched ECDSA verification (compressed 33-byte keys, 64-byte compact signatures), with
 * libsecp256k1 semantics: low-S required, r|s >= n is a parse error.
 * Replaces crypto/txscript/src/lib.rs:618-628. */
int kgv_ecdsa_verify(kgv_ctx* ctx, const uint8_t* pk33, const uint8_t* msg32, const uint8_t* sig64, size_t n, uint8_t* status);

/* Pack verdicts into a validity bitmap: bit i (LSB-first within each byte) = (status[i] == KGV_SIG_VALID).
 * bitmap has (n+7)/8 bytes.  This is the per-shard payload of the multi-GPU all-gather. */
int kgv_status_to_bitmap(kgv_ctx* ctx, const uint8_t* status, size_t n, uint8_t* bitmap);

/* ------------------------------------------------------------------------------------------------
 * Flat SoA transaction batches (little-endian, fixed-size records + one byte arena).
 * Mirror of the reference data model: Transaction consensus/core/src/tx.rs:165-185,
 * TransactionInput :93-101, TransactionOutpoint :72-77, TransactionOutput :121-125, UtxoEntry :49-57,
 * ScriptPublicKey tx/script_public_key.rs:22-25.  Offsets point into `bytes`.
 * ------------------------------------------------------------------------------------------------ */
typedef struct {
  uint32_t first_input, n_inputs;   /* range in inputs[]  */
  uint32_t first_output, n_outputs; /* range in outputs[] */
  uint64_t lock_time, gas;
  uint64_t mass;                    /* committed storage mass (tx.mass()) */
  uint32_t payload_off, payload_len;
  uint16_t version;
  uint8_t subnetwork_id[20];
  uint8_t flags;                    /* bit 0: coinbase (informational; derived from subnetwork_id) */
  uint8_t pad_;
} kgv_tx; /* 72 bytes */
typedef struct {
  uint8_t prev_txid[32];
  uint32_t prev_index;
  uint32_t sigscript_off, sigscript_len;
  uint8_t sig_op_count;
  uint8_t pad_[3];
  uint64_t sequence;
} kgv_input; /* 56 bytes */
typedef struct {
  uint64_t value;
  uint32_t script_off, script_len;
  uint16_t spk_version;
  uint8_t pad_[6];
} kgv_output; /* 24 bytes */
typedef struct {
  uint64_t amount;
  uint64_t block_daa_score;
  uint32_t script_off, script_len;  /* script_public_key.script */
  uint16_t spk_version;
  uint8_t is_coinbase;
  uint8_t pad_[5];                  /* pad_[0] != 0 in a populated batch marks the entry ABSENT (-> MissingTxOutpoints) */
} kgv_utxo_entry; /* 32 bytes */
typedef struct {
  const kgv_tx* txs; size_t n_txs;
  const kgv_input* inputs; size_t n_inputs;
  const kgv_output* outputs; size_t n_outputs;
  const kgv_utxo_entry* entries; /* one populated entry per input (PopulatedTransaction), or NULL */
  const uint8_t* bytes; size_t n_bytes;
} kgv_tx_batch; /* the arrays are all host pointers or all device pointers */

/* Transaction ids / hashes of every tx of the batch: out32 = n_txs * 32 bytes.
 * Replaces consensus/core/src/hashing/tx.rs:16-42 (`hash`, `id`; keyed BLAKE2b). */
int kgv_tx_ids(kgv_ctx* ctx, const kgv_tx_batch* batch, uint8_t* out32);
int kgv_tx_hashes(kgv_ctx* ctx, const kgv_tx_batch* batch, uint8_t* out32);

/* Signature hashes.  One item per signature check: `input` is the ABSOLUTE index into
 * batch->inputs / batch->entries, hash_type one of 01 02 04 81 82 84 (anything else yields an
 * all-0xFF digest, sighash_type.rs:50-56 InvalidSigHashType), ecdsa != 0 adds the
 * SHA-256 wrap.  Replaces calc_schnorr_signature_hash / calc_ecdsa_signature_hash
 * (consensus/core/src/hashing/sighash.rs:238-277); the five per-tx sub-hashes are computed once
 * per transaction (SigHashReusedValues, sighash.rs:14-138). */
typedef struct {
  uint32_t tx;
  uint32_t input;
  uint8_t hash_type;
  uint8_t ecdsa;
  uint8_t pad_[2];
} kgv_sighash_item; /* 12 bytes */
int kgv_sighash(kgv_ctx* ctx, const kgv_tx_batch* batch, const kgv_sighash_item* items, size_t n_items, uint8_t* out32);

/* ------------------------------------------------------------------------------------------------
 * UTXO-context validation (the rayon fan-out of utxo_validation.rs:262-338 as a batch call)
 * ------------------------------------------------------------------------------------------------ */
/* per-transaction verdicts = TxRuleError classes (consensus/core/src/errors/tx.rs:8-103) reachable from
 * validate_transaction_in_utxo_context (utxo_validation.rs:312-338) */
#define KGV_TX_OK 0
#define KGV_TX_MISSING_OUTPOINTS 1      /* MissingTxOutpoints                */
#define KGV_TX_IMMATURE_COINBASE 2      /* ImmatureCoinbaseSpend             */
#define KGV_TX_INPUT_AMOUNT_OVERFLOW 3  /* InputAmountOverflow               */
#define KGV_TX_INPUT_AMOUNT_TOO_HIGH 4  /* InputAmountTooHigh                */
#define KGV_TX_SPEND_TOO_HIGH 5         /* SpendTooHigh                      */
#define KGV_TX_MASS_INCOMPUTABLE 6      /* MassIncomputable                  */
#define KGV_TX_WRONG_MASS 7             /* WrongMass                         */
#define KGV_TX_SEQUENCE_LOCK 8          /* SequenceLockConditionsAreNotMet   */
#define KGV_TX_SIGNATURE_INVALID 9      /* SignatureInvalid(script_err)      */
#define KGV_TX_SIGNATURE_EMPTY 10       /* SignatureEmpty(script_err)        */
#define KGV_TX_NEEDS_HOST_VM 11         /* an input is not one of the GPU fast-path script classes: the host
                                           script engine must decide this transaction (all context checks passed) */
#define KGV_TX_SKIPPED_COINBASE 12      /* coinbase transactions are skipped (utxo_validation.rs:273) */
/* script errors = TxScriptError variants the standard classes can produce (crypto/txscript/errors) */
#define KGV_SCRIPT_OK 0
#define KGV_SCRIPT_EVAL_FALSE 1
#define KGV_SCRIPT_NULL_FAIL 2
#define KGV_SCRIPT_INVALID_SIGNATURE 3
#define KGV_SCRIPT_SIG_LENGTH 4
#define KGV_SCRIPT_PUBKEY_FORMAT 5
#define KGV_SCRIPT_INVALID_SIGHASH_TYPE 6
#define KGV_SCRIPT_EXCEEDED_SIGOP_LIMIT 7
#define KGV_SCRIPT_NONSTANDARD 255
/* TxValidationFlags (tx_validation_in_utxo_context.rs:20-31) */
#define KGV_FLAGS_FULL 0
#define KGV_FLAGS_SKIP_SCRIPT_CHECKS 1
#define KGV_FLAGS_SKIP_MASS_CHECK 2
/* extension: run ONLY check_scripts on populated entries (no maturity / amount / mass / sequence-lock rules).
 * Signatures are context free given the spent output (the sighash reads only the entry's
 * script_public_key and amount, sighash.rs:252-255), so a window of future blocks can be script-checked in one
 * large batch before the in-order pass, which then runs with KGV_FLAGS_SKIP_SCRIPT_CHECKS. */
#define KGV_FLAGS_SCRIPTS_ONLY 3

typedef struct {
  uint64_t coinbase_maturity;       /* Params::coinbase_maturity      */
  uint64_t storage_mass_parameter;  /* Params::storage_mass_parameter */
  uint64_t max_sompi;               /* constants::MAX_SOMPI           */
} kgv_params;
typedef struct {
  uint64_t fee;        /* calculated_fee (valid when status == KGV_TX_OK) */
  uint32_t fail_input; /* first failing input (index within the tx) for status 2, 9, 10, 11 */
  uint8_t status;      /* KGV_TX_*     */
  uint8_t script_err;  /* KGV_SCRIPT_* */
  uint8_t pad_[2];
} kgv_tx_result; /* 16 bytes */

/* validate_populated_transaction_and_get_fee for every tx of a batch whose entries are already
 * populated (batch->entries != NULL) — tx_validation_in_utxo_context.rs:34-61 incl. check_scripts
 * (:157-196) for the standard script classes.  results: n_txs records. */
int kgv_validate_populated(kgv_ctx* ctx, const kgv_tx_batch* batch, uint64_t pov_daa_score, uint32_t flags, const kgv_params* params,
                           kgv_tx_result* results);

/* GPU-resident UTXO set: open-addressed hash table keyed by the 36-byte outpoint (txid || index LE),
 * 128-byte slots (entry + up to 68 script bytes inline, longer scripts in an overflow arena).
 * It plays the role of the base layer of every composed view: DbUtxoSetStore
 * (consensus/src/model/stores/utxo_set.rs:31-44,107-112,143-152) and UtxoCollection
 * (consensus/core/src/utxo/utxo_collection.rs:5,28-32). */
typedef struct kgv_utxo_table kgv_utxo_table;
int kgv_utxo_create(kgv_ctx* ctx, uint64_t capacity_slots /* rounded up to a power of two */, kgv_utxo_table** out);
void kgv_utxo_destroy(kgv_ctx* ctx, kgv_utxo_table* t);
/* UtxoView::get for n outpoints: found[i] in {0,1}; entries[i].script_off = i*script_stride into scripts_out
 * (scripts longer than script_stride are truncated there; script_len is always the true length). */
int kgv_utxo_lookup(kgv_ctx* ctx, kgv_utxo_table* t, const uint8_t* keys36, size_t n, kgv_utxo_entry* entries, uint8_t* scripts_out,
                    uint32_t script_stride, uint8_t* found);
/* write_diff_batch (utxo_set.rs:107-112): delete the removed outpoints, then put the added ones.
 * add_entries[i].script_off/len point into add_bytes.  Keys of one call must be distinct.
 * rem_status[i]: 1 = was present, 0 = absent; add_status[i]: 1 = inserted, 2 = replaced an existing entry. */
int kgv_utxo_apply_diff(kgv_ctx* ctx, kgv_utxo_table* t, const uint8_t* rem_keys36, size_t n_rem, uint8_t* rem_status,
                        const uint8_t* add_keys36, const kgv_utxo_entry* add_entries, const uint8_t* add_bytes, size_t n_add_bytes, size_t n_add,
                        uint8_t* add_status);
/* number of live entries, and an order-independent digest of the set: the sum modulo 2^256 of the keyed
 * BLAKE2b "MuHashElement" hashes of every (outpoint, entry) (consensus/core/src/muhash.rs:47-59). */
int kgv_utxo_count(kgv_ctx* ctx, kgv_utxo_table* t, uint64_t* count);
int kgv_utxo_digest(kgv_ctx* ctx, kgv_utxo_table* t, uint8_t out32[32]);
/* DbUtxoSetStore::iterator (utxo_set.rs:114-129): every live entry as (key, entry record, script bytes), in the table's (arbitrary) order -
 * what the syncer side of a pruning-point import streams out, and what `virtual.utxo_set := pruning-point utxo_set` copies
 * (consensus/src/pipeline/virtual_processor/processor.rs:1150-1158).  Arrays all host or all device.  With keys36 == NULL only the sizes are
 * returned (*n_out entries, *bytes_out script bytes); arrays smaller than that give KGV_ERR_NOMEM with the sizes still set. */
int kgv_utxo_export(kgv_ctx* ctx, kgv_utxo_table* t, uint8_t* keys36, kgv_utxo_entry* entries, uint8_t* bytes, size_t max_n, size_t bytes_cap, size_t* n_out,
                    size_t* bytes_out);
/* append_imported_pruning_point_utxos (consensus/src/consensus/mod.rs:1070-1083) for one chunk of the imported UTXO set: the entries are
 * written into the table (write_many) and MuHash::from_utxo of every (outpoint, entry), reduced, is combined into the running multiset
 * numerator384 (host value, in / out, 384 little-endian bytes; start from 1).  The caller then compares kgv_muhash_finalize(numerator, 1)
 * with the new pruning point's header.utxo_commitment (processor.rs:1133-1139: ImportedMultisetHashMismatch) and validates the pruning
 * point's own transactions with kgv_validate_txs (:1162-1172). */
int kgv_utxo_import_chunk(kgv_ctx* ctx, kgv_utxo_table* t, const uint8_t* keys36, const kgv_utxo_entry* entries, const uint8_t* bytes, size_t n_bytes, size_t n,
                          uint8_t* numerator384);

/* Composed views (consensus/core/src/utxo/utxo_view.rs:22-35 ComposedUtxoView / UtxoViewComposition::compose; UtxoDiff utxo_diff.rs:15-19):
 * a DIFF LAYER on the device.  The returned handle is a kgv_utxo_table that every call accepting a table accepts; it behaves as base ∘ diff:
 *   lookups (kgv_utxo_lookup, the populate step of kgv_validate_txs / kgv_replay_window) probe the layer first - an entry it added is found, an
 *   outpoint it removed is absent - and fall through to `base` otherwise (`base` may itself be a view: views nest like the reference's
 *   utxo_set ∘ accumulated_diff ∘ mergeset_diff, processor.rs:437,527);
 *   writes (kgv_utxo_apply_diff, kgv_utxo_apply_accepted, the in-order pass of kgv_replay_window) go to the layer only: removing an entry that
 *   lives below records a removal marker, removing the layer's own addition cancels it, re-adding a removed outpoint keeps the lower entry hidden.
 * `base` is never modified until kgv_utxo_view_commit folds the layer into it (write_diff_batch, utxo_set.rs:107-112); kgv_utxo_view_discard drops
 * the layer's content (a candidate chain that lost).  Count / digest / MuHash are defined on plain tables only. */
int kgv_utxo_view_create(kgv_ctx* ctx, kgv_utxo_table* base, uint64_t capacity_slots, kgv_utxo_table** out);
int kgv_utxo_view_commit(kgv_ctx* ctx, kgv_utxo_table* view);
int kgv_utxo_view_discard(kgv_ctx* ctx, kgv_utxo_table* view);

/* validate_transactions_in_parallel (utxo_validation.rs:262-278) against the table: populate every input
 * by table lookup (:319-327), then as kgv_validate_populated.  batch->entries is ignored.
 * Unlike kgv_validate_populated this call never reports KGV_TX_NEEDS_HOST_VM: transactions with non-standard scripts are
 * decided inside the call by the host script engine on the entries the table returned (the reference accepts ANY
 * transaction whose scripts execute successfully, utxo_validation.rs:282-309). */
int kgv_validate_txs(kgv_ctx* ctx, kgv_utxo_table* t, const kgv_tx_batch* batch, uint64_t pov_daa_score, uint32_t flags, const kgv_params* params,
                     kgv_tx_result* results);
/* UtxoDiff::add_transaction (utxo_diff.rs:233-247) applied directly to the table for every tx with
 * accept[i] != 0: spent outpoints are erased, outputs inserted with block_daa_score = pov_daa_score and
 * is_coinbase of the tx (tx ids are computed on the device). */
int kgv_utxo_apply_accepted(kgv_ctx* ctx, kgv_utxo_table* t, const kgv_tx_batch* batch, const uint8_t* accept, uint64_t pov_daa_score);

/* ------------------------------------------------------------------------------------------------
 * SigCache: Cache<SigCacheKey, bool> (crypto/txscript/src/caches.rs:14-55; consulted at crypto/txscript/src/lib.rs:589-603, 624-638;
 * created with 10 000 entries and shared by every clone of the TransactionValidator, transaction_validator/mod.rs:48).
 * A device-resident, bounded table of verdicts keyed by BLAKE2b-256(kind || signature || public key || message).  Attached to a context
 * (kgv_set_sigcache; several contexts of one device may share one cache, as the clones share the Arc), it is consulted by the script
 * phase of kgv_validate_populated / kgv_validate_txs / kgv_replay_window: pairs seen before are answered from the table, only the
 * misses reach the verification kernels, and their verdicts (true AND false; never parse errors, which the reference raises before its
 * cache) are remembered.  Full neighbourhoods evict a pseudo-random entry (caches.rs:49-51).  Results never change, only speed:
 * block-template building and block validation re-meet what the mempool verified (processor.rs:853-914).
 * capacity is rounded up to a power of two.  counters: hits = get_counts, inserts = insert_counts (caches.rs:57-93).
 * ------------------------------------------------------------------------------------------------ */
typedef struct kgv_sigcache kgv_sigcache;
int kgv_sigcache_create(kgv_ctx* ctx, uint64_t capacity, kgv_sigcache** out);
void kgv_sigcache_destroy(kgv_sigcache* cache);
int kgv_sigcache_clear(kgv_ctx* ctx, kgv_sigcache* cache);
int kgv_sigcache_counters(kgv_ctx* ctx, kgv_sigcache* cache, uint64_t* hits, uint64_t* inserts, uint64_t* lookups, uint64_t* evictions);
int kgv_set_sigcache(kgv_ctx* ctx, kgv_sigcache* cache /* NULL: detach */);

/* ------------------------------------------------------------------------------------------------
 * Multi-GPU (SURVEY.md §8b kgv_shard_allgather, §8e): signature batches shard across GPUs as contiguous ranges, one context (and
 * normally one process) per GPU; the only exchange step of the path is "every rank ends up with every shard's verdicts".
 * The reference has no counterpart (rayon on one host, utxo_validation.rs:269-277).
 *
 * A communicator binds a context to its place among n_ranks and offers two transports:
 *   NCCL  (id != NULL): kgv_shard_allgather = ncclAllGather on the context's stream.  libnccl.so.2 is resolved at run time;
 *         KGV_ERR_NCCL if it is missing.  The 128-byte id comes from kgv_comm_unique_id on rank 0 and reaches the other ranks
 *         through whatever the host uses (MPI, TCP, torch.distributed).
 *   peer  (slice_capacity_bytes != 0): every rank owns receive buffers its peers map (kgv_comm_export -> host exchanges the
 *         handles -> kgv_comm_import; or kgv_comm_connect_local for several contexts of ONE process).  kgv_shard_publish_* is the
 *         PRODUCING kernel writing its shard straight into every peer over NVLink and raising an epoch flag there;
 *         kgv_shard_wait waits on local flags.  No collective, no host rendezvous.  Every rank must publish the same sequence
 *         of epochs and wait for epoch e before publishing e + 1.
 * ------------------------------------------------------------------------------------------------ */
typedef struct kgv_comm kgv_comm;
#define KGV_COMM_ID_BYTES 128
#define KGV_COMM_HANDLE_BYTES 64
int kgv_comm_unique_id(uint8_t id[KGV_COMM_ID_BYTES]);
int kgv_comm_create(kgv_ctx* ctx, int n_ranks, int rank, const uint8_t* id /* NULL: no NCCL */, size_t slice_capacity_bytes /* 0: no peer buffers */,
                    kgv_comm** out);
void kgv_comm_destroy(kgv_comm* comm);
int kgv_comm_export(kgv_comm* comm, uint8_t handle[KGV_COMM_HANDLE_BYTES]);
int kgv_comm_import(kgv_comm* comm, const uint8_t* handles /* n_ranks * KGV_COMM_HANDLE_BYTES, indexed by rank */);
int kgv_comm_connect_local(kgv_comm* const* comms, int n /* comms[i] has rank i */);
/* all_shards[r * nbytes_per_rank ..) = rank r's local_shard, on every rank (device pointers; in place allowed as in NCCL). */
int kgv_shard_allgather(kgv_ctx* ctx, kgv_comm* comm, const uint8_t* local_shard, size_t nbytes_per_rank, uint8_t* all_shards);
/* peer transport.  publish_bitmap: kgv_status_to_bitmap fused with the transfer (status: n device bytes; the shard is 4*ceil(n/32) bytes);
 * publish_bytes: a raw device array.  *epoch_out names the exchange.  wait: enqueue the wait for `epoch` and (all_shards != NULL)
 * copy the n_ranks received shards, nbytes_per_rank each, into one contiguous device array. */
int kgv_shard_publish_bitmap(kgv_ctx* ctx, kgv_comm* comm, const uint8_t* status, size_t n, uint64_t* epoch_out);
int kgv_shard_publish_bytes(kgv_ctx* ctx, kgv_comm* comm, const uint8_t* src, size_t nbytes, uint64_t* epoch_out);
int kgv_shard_wait(kgv_ctx* ctx, kgv_comm* comm, uint64_t epoch, size_t nbytes_per_rank, uint8_t* all_shards);
/* Shard the signature checks of this context's validation calls (kgv_replay_window, kgv_validate_*) over the communicator's
 * ranks (BASELINE configs[4]): every rank runs the same call on the same batch against its own replica of the UTXO table;
 * the candidate (signature, key) pairs are split into n_ranks contiguous ranges, each rank verifies one range and the verdicts
 * are exchanged (peer transport if connected, else NCCL) before the scripts are resolved - identically on every rank.
 * comm == NULL switches sharding off. */
int kgv_set_sharding(kgv_ctx* ctx, kgv_comm* comm);

/* ------------------------------------------------------------------------------------------------
 * DAG replay: calculate_utxo_state / verify_expected_utxo_state (utxo_validation.rs:110-173,182-228) for a WINDOW of
 * blocks as one device-resident call; the loop simpa times (simpa/src/main.rs:454-460).
 *
 * The batch holds the transactions of all blocks of the window, block after block; blocks[] (a HOST array) tiles it.
 * Transaction 0 of every block is its coinbase and is skipped by position (utxo_validation.rs:273).  Semantics are those
 * of processing the blocks one by one, in order, against the table:
 *     validate_transactions_in_parallel(block txs, table, pov_daa_score, Full | SkipScriptChecks)
 *     UtxoDiff::add_transaction for every accepted transaction (and for the coinbase when ACCEPT_COINBASE is set)
 * but every signature of the window is verified in one batch up front (signatures are context free given the spent
 * output) and the in-order pass is a single persistent kernel.
 * All merged blocks of one chain block carry that chain block's pov_daa_score (:124-151).
 * ------------------------------------------------------------------------------------------------ */
#define KGV_REPLAY_ACCEPT_COINBASE 1u /* the block is the selected parent of its chain block: its coinbase is accepted (:116-121) */
#define KGV_REPLAY_SKIP_SCRIPTS 2u    /* TxValidationFlags::SkipScriptChecks for this block (selected parent, :138-140) */
#define KGV_REPLAY_VERIFY_ONLY 4u     /* validate against the current view, apply nothing (verify_expected_utxo_state, :219-225) */
typedef struct {
  uint32_t first_tx, n_txs; /* range of batch->txs; tx first_tx is the block's coinbase */
  uint64_t pov_daa_score;
  uint32_t flags;           /* KGV_REPLAY_* */
  uint32_t pad_;
} kgv_replay_block; /* 24 bytes */
typedef struct {
  uint64_t n_accepted;   /* accepted non-coinbase transactions */
  uint64_t n_sig_checks; /* candidate (signature, key) pairs verified in the pre-check */
  uint64_t n_host_vm;    /* transactions decided by the host script engine */
  float pre_check_ms;    /* device time of the batched script pre-check (tx ids, window map, populate, sighash, verify, resolve) */
  float in_order_ms;     /* device time of the in-order pass */
} kgv_replay_stats;
/* results: n_txs records (host or device memory): the UTXO-context verdict when the context rules fail, else the script
 * verdict (KGV_TX_SKIPPED_COINBASE for coinbases).  accept (may be NULL): n_txs bytes, 1 = folded into the table.
 * stats may be NULL. */
int kgv_replay_window(kgv_ctx* ctx, kgv_utxo_table* t, const kgv_tx_batch* batch, const kgv_replay_block* blocks, size_t n_blocks,
                      const kgv_params* params, kgv_tx_result* results, uint8_t* accept, kgv_replay_stats* stats);

/* The multiset of what the LAST kgv_replay_window call of this context accepted, per group of blocks: group g = blocks
 * [group_first_block[g], group_first_block[g+1]) of that window (the mergeset of one chain block); values768[g] = (numerator || denominator) of
 * MuHash::from_transaction over the accepted transactions of the group incl. accepted coinbases (utxo_validation.rs:116-121,144; spent entries as
 * found at each block's position).  Must directly follow that kgv_replay_window call (no other batch call in between).
 * With kgv_muhash_prefix_combine and kgv_muhash_finalize_batch this yields every chain block's utxo_commitment of a window (:188-192). */
int kgv_replay_muhash(kgv_ctx* ctx, const uint32_t* group_first_block, size_t n_groups, uint8_t* values768);
/* Overlap of the next window's upload with the current window's compute (IBD: the caller knows the blocks ahead).  A HOST batch is checked and
 * copied to the device on a side stream into a second staging buffer; the next kgv_replay_window / kgv_validate_txs / kgv_tx_ids ... call that is
 * handed exactly this batch (the same arrays and sizes) takes that buffer over instead of uploading.  The arrays must stay unchanged - and be
 * page-locked for the copy to really be asynchronous - until that call.  A device-resident batch is a no-op.  Two batches can be in flight (the
 * usual order is: prefetch window i+1, then the synchronous call for window i, whose own copy was prefetched one step earlier); a third
 * prefetch replaces the older one. */
int kgv_batch_prefetch(kgv_ctx* ctx, const kgv_tx_batch* batch);

/* ------------------------------------------------------------------------------------------------
 * Merkle roots: crypto/merkle/src/lib.rs:3-30 calc_merkle_root / merkle_hash.
 * ------------------------------------------------------------------------------------------------ */
/* n_groups independent trees over one flattened array of 32-byte hashes: group g = hashes [first[g], first[g+1]).
 * `first` (n_groups + 1 offsets) is a HOST array; hashes32 / roots32 are both host or both device.  An empty group
 * yields ZERO_HASH, a single hash is its own root.  Serves hash_merkle_root (with kgv_tx_hashes) and
 * calc_accepted_id_merkle_root's inner root (utxo_validation.rs:401-410, with kgv_tx_ids of the accepted txs). */
int kgv_merkle_roots(kgv_ctx* ctx, const uint8_t* hashes32, const uint32_t* first, uint32_t n_groups, uint8_t* roots32);
/* calc_hash_merkle_root (consensus/core/src/merkle.rs:5-7; checked by body_validation_in_isolation.rs:34-40) for every
 * block of a batch: block b = transactions [block_first_tx[b], block_first_tx[b+1]) (host array of n_blocks + 1). */
int kgv_block_hash_merkle_roots(kgv_ctx* ctx, const kgv_tx_batch* batch, const uint32_t* block_first_tx, uint32_t n_blocks, uint8_t* roots32);

/* Block-body set checks of validate_body_in_isolation (body_validation_in_isolation.rs:13-23,95-131) for many blocks at
 * once: check_duplicate_transactions, check_block_double_spends, check_no_chained_transactions, in that order; the
 * reported item is the first offender in the reference's iteration order (the tx index for duplicates, the absolute
 * input index otherwise).  grid y dimension = n_blocks (<= 65535 per call). */
#define KGV_BLOCK_OK 0u
#define KGV_BLOCK_DUPLICATE_TRANSACTIONS 1u      /* RuleError::DuplicateTransactions(tx id of txs[index]) */
#define KGV_BLOCK_DOUBLE_SPEND_IN_SAME_BLOCK 2u  /* RuleError::DoubleSpendInSameBlock(inputs[index].previous_outpoint) */
#define KGV_BLOCK_CHAINED_TRANSACTION 3u         /* RuleError::ChainedTransaction(inputs[index].previous_outpoint) */
typedef struct kgv_block_check { uint32_t status, index; } kgv_block_check;
int kgv_block_set_checks(kgv_ctx* ctx, const kgv_tx_batch* batch, const uint32_t* block_first_tx, uint32_t n_blocks, kgv_block_check* out);

/* ------------------------------------------------------------------------------------------------
 * K8 MuHash: crypto/muhash/src/lib.rs, u3072.rs; consensus/core/src/muhash.rs.
 * A MuHash is the pair (numerator, denominator) of residues modulo 2^3072 - 1103717 (lib.rs:32-35); every function
 * here reads / writes them as 384 little-endian bytes each, CANONICAL (in [0, p)) on output.  The reference's
 * transient non-canonical representations (u3072.rs:49-57) are unobservable through serialize()/finalize().
 * Pointers of one call are all host or all device.
 * ------------------------------------------------------------------------------------------------ */
/* MuHash::add_element / remove_element (lib.rs:61-74) for n byte strings data[offsets[i] .. offsets[i+1]):
 * remove[i] != 0 multiplies the element into the denominator, else into the numerator (remove may be NULL).
 * Starts from the empty MuHash (1, 1). */
int kgv_muhash_elements(kgv_ctx* ctx, const uint8_t* data, const uint64_t* offsets, const uint8_t* remove, size_t n, uint8_t* numerator384,
                        uint8_t* denominator384);
/* The MuHash half of validate_transactions_with_muhash_in_parallel (utxo_validation.rs:282-309): MuHash::from_transaction
 * (consensus/core/src/muhash.rs:16-27,35-39) of every tx with accept[i] != 0, all combined.  Populated entries come
 * from batch->entries, or from `table` when it is non-NULL (call BEFORE kgv_utxo_apply_accepted erases them). */
int kgv_muhash_txs(kgv_ctx* ctx, kgv_utxo_table* table, const kgv_tx_batch* batch, const uint8_t* accept, uint64_t pov_daa_score,
                   uint8_t* numerator384, uint8_t* denominator384);
/* MuHash::combine (lib.rs:91-96): a.numerator *= b.numerator, a.denominator *= b.denominator */
int kgv_muhash_combine(kgv_ctx* ctx, uint8_t* numerator_a, uint8_t* denominator_a, const uint8_t* numerator_b, const uint8_t* denominator_b);
/* MuHash::serialize + finalize (lib.rs:98-115): serialized = numerator / denominator (0 has inverse 0, u3072.rs:163-165),
 * hash = BLAKE2b-256 keyed "MuHashFinalize".  Sequential by nature (one modular inversion: 3 072 dependent squarings);
 * the reference calls it once per chain block outside the parallel section.  serialized384 may be NULL. */
int kgv_muhash_finalize(kgv_ctx* ctx, const uint8_t* numerator384, const uint8_t* denominator384, uint8_t* serialized384, uint8_t* hash32);
/* n finalizations at once: hashes32[i] = MuHash{numerator_i, denominator_i}.finalize() (lib.rs:98-115); value i sits pitch_bytes after value
 * i - 1 (384 for plain arrays, 768 for (numerator || denominator) records).  ONE modular inversion for the whole batch (Montgomery's
 * trick over prefix / suffix products built by parallel scans): a chain block's commitment costs five multiplications instead of
 * 3 072 squarings.  serialized384 (n * 384 contiguous bytes) may be NULL. */
int kgv_muhash_finalize_batch(kgv_ctx* ctx, const uint8_t* numerators384, const uint8_t* denominators384, size_t n, size_t pitch_bytes, uint8_t* serialized384,
                              uint8_t* hashes32);
/* The MuHash::combine chain of a replay (utxo_validation.rs:144): values768 holds n (numerator || denominator) records; on return record i is
 * init * record 0 * ... * record i (canonical).  init768 may be NULL (= the empty MuHash). */
int kgv_muhash_prefix_combine(kgv_ctx* ctx, const uint8_t* init768, uint8_t* values768, size_t n);
/* MuHash::add_utxo (consensus/core/src/muhash.rs:28-33) over every live entry of the table: the UTXO-set commitment
 * numerator (denominator 1). */
int kgv_utxo_muhash(kgv_ctx* ctx, kgv_utxo_table* t, uint8_t* numerator384);

/* ------------------------------------------------------------------------------------------------
 * Host script engine for non-standard scripts (KGV_TX_NEEDS_HOST_VM)
 * Complete restatement of TxScriptEngine (crypto/txscript/src/lib.rs:83-98,276-643, opcodes/mod.rs,
 * data_stack.rs); signature checks are resolved by a verdict provider, never computed on the CPU.
 * ------------------------------------------------------------------------------------------------ */
/* further TxScriptError variants only the full engine can produce (crypto/txscript/errors/src/lib.rs) */
#define KGV_SCRIPT_NOT_PUSH_ONLY 8
#define KGV_SCRIPT_CLEAN_STACK 9
#define KGV_SCRIPT_EMPTY_STACK 10
#define KGV_SCRIPT_ELEMENT_TOO_BIG 11
#define KGV_SCRIPT_TOO_MANY_OPERATIONS 12
#define KGV_SCRIPT_STACK_SIZE_EXCEEDED 13
#define KGV_SCRIPT_OPCODE_DISABLED 14
#define KGV_SCRIPT_OPCODE_RESERVED 15
#define KGV_SCRIPT_INVALID_OPCODE 16
#define KGV_SCRIPT_MALFORMED_PUSH 17
#define KGV_SCRIPT_MALFORMED_PUSH_SIZE 18
#define KGV_SCRIPT_NOT_MINIMAL_DATA 19
#define KGV_SCRIPT_UNBALANCED_CONDITIONAL 20
#define KGV_SCRIPT_COND_STACK_EMPTY 21    /* InvalidState("condition stack empty")       */
#define KGV_SCRIPT_EXPECTED_BOOLEAN 22    /* InvalidState("expected boolean")            */
#define KGV_SCRIPT_PICK_INVALID 23        /* InvalidState("pick at an invalid location") */
#define KGV_SCRIPT_ROLL_INVALID 24        /* InvalidState("roll at an invalid location") */
#define KGV_SCRIPT_VERIFY 25
#define KGV_SCRIPT_EARLY_RETURN 26
#define KGV_SCRIPT_INVALID_STACK_OPERATION 27
#define KGV_SCRIPT_NUMBER_TOO_BIG 28
#define KGV_SCRIPT_INVALID_PUBKEY_COUNT 29
#define KGV_SCRIPT_INVALID_SIGNATURE_COUNT 30
#define KGV_SCRIPT_UNSATISFIED_LOCKTIME 31
#define KGV_SCRIPT_SCRIPT_SIZE 32
#define KGV_SCRIPT_NO_SCRIPTS 33
#define KGV_SCRIPT_INVALID_INPUT_INDEX 34
#define KGV_SCRIPT_INVALID_OUTPUT_INDEX 35
#define KGV_SCRIPT_SERIALIZATION 36
#define KGV_SCRIPT_NEEDS_SIG_VERDICTS 254 /* kgv_script_execute only: the provider did not know a verdict */

typedef struct {
  uint32_t tx, input;      /* absolute input index into batch->inputs */
  uint8_t hash_type, ecdsa;
  uint8_t key_len;         /* 32 or 33 */
  uint8_t pad_;
  uint8_t key[33];
  uint8_t sig[64];
  uint8_t pad2_[3];
} kgv_sig_request; /* 112 bytes */
/* returns KGV_SIG_* (0..3), or a negative value if the verdict is not available */
typedef int (*kgv_verdict_fn)(void* user, const kgv_sig_request* request);

/* TxScriptEngine::from_transaction_input(..).execute() for ONE input of a HOST-resident populated batch
 * (lib.rs:276-300,399-449); `input_index` is relative to the transaction.  No GPU context is involved:
 * every check_schnorr/ecdsa_signature (lib.rs:574-643) asks `verdict` (a SigCache, the GPU batch, ...).
 * *script_err receives a KGV_SCRIPT_* code (0 = the input's scripts succeed). */
int kgv_script_execute(const kgv_tx_batch* batch, uint32_t tx, uint32_t input_index, kgv_verdict_fn verdict, void* user, uint8_t* script_err);

/* check_scripts (tx_validation_in_utxo_context.rs:162-200) with the full host engine for the listed
 * transactions of a HOST-resident populated batch; the signature checks the scripts reach are gathered,
 * hashed and verified on the GPU in batches (as many rounds as the scripts' control flow needs).
 * results[i] belongs to tx_indices[i]: status KGV_TX_OK / KGV_TX_SIGNATURE_INVALID / KGV_TX_SIGNATURE_EMPTY. */
int kgv_check_scripts_host(kgv_ctx* ctx, const kgv_tx_batch* batch, const uint32_t* tx_indices, size_t n, kgv_tx_result* results);

/* ------------------------------------------------------------------------------------------------
 * Persistence formats either side of the path: the RocksDB rows of DbUtxoSetStore.  Host functions (the store lives on
 * the host): what a shim runs between the database and kgv_utxo_apply_diff / kgv_utxo_lookup (write_diff_batch utxo_set.rs:107-112, the
 * iterator feeding the pruning-point UTXO-set import processor.rs:1126-1200).
 *   key row    txid(32) || index u32 LE with trailing zero bytes trimmed, at least one kept (UtxoKey, utxo_set.rs:31-62) - WITHOUT the store's
 *              prefix bytes, which the caller prepends
 *   value row  bincode::serialize(&UtxoEntry) (database/src/access.rs:139): amount u64 || spk version u16 || script length u64 || script ||
 *              block_daa_score u64 || is_coinbase u8, all little-endian
 * Rows are packed back to back; *_off has n + 1 entries.  encode: key_rows / value_rows may be NULL (with capacity 0) to size the buffers first.
 * decode: scripts are appended to bytes_out and entries[i].script_off points there; malformed rows -> KGV_ERR_ARG.
 * ------------------------------------------------------------------------------------------------ */
int kgv_utxo_rows_encode(const uint8_t* keys36, const kgv_utxo_entry* entries, const uint8_t* bytes, size_t n_bytes, size_t n, uint8_t* key_rows, uint64_t* key_off,
                         uint8_t* value_rows, uint64_t* value_off, size_t key_cap, size_t value_cap);
int kgv_utxo_rows_decode(const uint8_t* key_rows, const uint64_t* key_off, const uint8_t* value_rows, const uint64_t* value_off, size_t n, uint8_t* keys36,
                         kgv_utxo_entry* entries, uint8_t* bytes_out, size_t bytes_cap, size_t* bytes_used);

/* Test / audit hook: affine coordinates (x||y, 32-byte big-endian each) of entry v (1..65535) of
 * generator table `which` (0: v*G, 1: v*2^128*G) as built on the device. */
int kgv_gtable_entry(kgv_ctx* ctx, int which, uint32_t v, uint8_t out_xy[64]);

/* Test / audit hook: verifies ONE Schnorr triple (host pointers) on the device and returns the
 * traced intermediates: trace_words[stage*16 + i], KGV_TRACE_STAGES stages of 16 u32 words
 * (stage numbering in csrc/kgv_verify.cuh).  tests/ compare it with the host build of the same
 * device code to localise a divergence. */
#define KGV_TRACE_STAGES 32
int kgv_debug_schnorr_trace(kgv_ctx* ctx, const uint8_t* pk32, const uint8_t* msg32, const uint8_t* sig64, uint32_t* trace_words,
                            uint8_t* status);

/* Test / audit hook: runs one arithmetic primitive (its PTX body) on n operand pairs on the device.
 * in_words / out_words: n x 16 u32 (a[8] || b[8] little-endian limbs in; result limbs out).
 * op: 0 mul_wide 1 sqr_wide 2 fe_mul 3 fe_sqr 4 sc_mul 5 sc_sqr 6 sc_inv 7 fe_inv 8 fe_add 9 fe_sub
 *     10 mul_wide+reduce mod n  11 glv_split. */
int kgv_debug_selftest(kgv_ctx* ctx, int op, const uint32_t* in_words, uint32_t* out_words, size_t n);

#ifdef __cplusplus
}
#endif
#endif /* KGV_H */
