// kgv_script_vm.h — host-side script engine for the inputs the GPU fast path declines
// (KGV_TX_NEEDS_HOST_VM): a complete restatement of the reference's txscript engine
//   crypto/txscript/src/lib.rs:83-98,276-470      TxScriptEngine (execute, execute_script, execute_opcode)
//   crypto/txscript/src/lib.rs:474-643            multisig, check_schnorr/ecdsa_signature
//   crypto/txscript/src/opcodes/mod.rs            all 256 opcodes, minimal-push rule, disabled / reserved lists
//   crypto/txscript/src/opcodes/macros.rs:1-42    push-opcode parsing (MalformedPush / MalformedPushSize)
//   crypto/txscript/src/data_stack.rs:87-330      script numbers, booleans, stack primitives
//   crypto/txscript/src/runtime_sig_op_counter.rs sig-op budget
// Signature checks are NOT computed here: the engine asks a verdict provider (backed by the GPU batch
// verifier); an unknown verdict suspends the run (SERR_NEEDS_SIG_VERDICTS) until the batch is verified.
#pragma once
#include <cstddef>
#include <cstdint>
#include <functional>
#include <vector>

#include "../../../include/kgv.h"

namespace kgv_host {

// TxScriptError variants (crypto/txscript/errors/src/lib.rs); 0..7 and 255 coincide with KGV_SCRIPT_* of include/kgv.h
enum ScriptErr : uint8_t {
  SERR_OK = 0, SERR_EVAL_FALSE = 1, SERR_NULL_FAIL = 2, SERR_INVALID_SIGNATURE = 3, SERR_SIG_LENGTH = 4, SERR_PUBKEY_FORMAT = 5,
  SERR_INVALID_SIGHASH_TYPE = 6, SERR_EXCEEDED_SIGOP_LIMIT = 7, SERR_NOT_PUSH_ONLY = 8, SERR_CLEAN_STACK = 9, SERR_EMPTY_STACK = 10,
  SERR_ELEMENT_TOO_BIG = 11, SERR_TOO_MANY_OPERATIONS = 12, SERR_STACK_SIZE_EXCEEDED = 13, SERR_OPCODE_DISABLED = 14, SERR_OPCODE_RESERVED = 15,
  SERR_INVALID_OPCODE = 16, SERR_MALFORMED_PUSH = 17, SERR_MALFORMED_PUSH_SIZE = 18, SERR_NOT_MINIMAL_DATA = 19, SERR_UNBALANCED_CONDITIONAL = 20,
  SERR_COND_STACK_EMPTY = 21 /* InvalidState("condition stack empty") */, SERR_EXPECTED_BOOLEAN = 22 /* InvalidState("expected boolean") */,
  SERR_PICK_INVALID = 23, SERR_ROLL_INVALID = 24, SERR_VERIFY = 25, SERR_EARLY_RETURN = 26, SERR_INVALID_STACK_OPERATION = 27,
  SERR_NUMBER_TOO_BIG = 28, SERR_INVALID_PUBKEY_COUNT = 29, SERR_INVALID_SIGNATURE_COUNT = 30, SERR_UNSATISFIED_LOCKTIME = 31, SERR_SCRIPT_SIZE = 32,
  SERR_NO_SCRIPTS = 33, SERR_INVALID_INPUT_INDEX = 34, SERR_INVALID_OUTPUT_INDEX = 35, SERR_SERIALIZATION = 36,
  SERR_NEEDS_SIG_VERDICTS = 254, SERR_NONSTANDARD = 255
};

struct SigRequest {  // one (signature, key) check of one input
  uint32_t tx, input_abs;
  uint8_t hash_type, ecdsa;
  std::vector<uint8_t> key;  // 32 or 33 bytes
  uint8_t sig[64];
};
// returns KGV_SIG_* (0..3) or -1 when the verdict is not known yet
using VerdictFn = std::function<int(const SigRequest&)>;

// Runs the engine on input `input_index` (relative to the tx) of transaction `tx` of a HOST-resident,
// populated batch.  When a verdict is missing the request is appended to `missing` and
// SERR_NEEDS_SIG_VERDICTS is returned.
ScriptErr execute_input(const kgv_tx_batch& b, uint32_t tx, uint32_t input_index, const VerdictFn& verdict, std::vector<SigRequest>* missing);

}  // namespace kgv_host
