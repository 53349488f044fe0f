// Exercises the C++ host mirror (include/kgv.hpp) on data dumped by tests/test_gpu_cpp_mirror.py and prints results as
// plain text for the Python side to compare with the oracle / the reference's known answers.  Built by
// __graft_entry__.build() (g++, links libkgv.so); needs a GPU to RUN.
//   host_mirror_test <dir>
// <dir> holds txs.bin inputs.bin outputs.bin entries.bin arena.bin (flat records of include/kgv.h), blocks.bin (u32 offsets),
// fund_keys.bin fund_entries.bin fund_arena.bin (the UTXO entries the batch spends), elements.txt (hex lines; "-" prefix = remove),
// triples.bin (n x (32 pk, 32 msg, 64 sig)).
#include <algorithm>
#include <cstdio>
#include <fstream>
#include <iostream>
#include <sstream>

#include "../../include/kgv.hpp"

template <class T>
static std::vector<T> slurp(const std::string& path) {
  std::ifstream f(path, std::ios::binary);
  if (!f) throw std::runtime_error("cannot open " + path);
  std::vector<char> raw((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
  std::vector<T> v(raw.size() / sizeof(T));
  std::memcpy(v.data(), raw.data(), v.size() * sizeof(T));
  return v;
}
static std::string hex(const uint8_t* p, size_t n) {
  static const char* d = "0123456789abcdef";
  std::string s;
  for (size_t i = 0; i < n; i++) { s.push_back(d[p[i] >> 4]); s.push_back(d[p[i] & 15]); }
  return s;
}
static std::vector<uint8_t> unhex(const std::string& s) {
  std::vector<uint8_t> v(s.size() / 2);
  for (size_t i = 0; i < v.size(); i++) v[i] = (uint8_t)std::stoi(s.substr(2 * i, 2), nullptr, 16);
  return v;
}

int main(int argc, char** argv) {
  if (argc < 2) { std::fprintf(stderr, "usage: %s <dir>\n", argv[0]); return 2; }
  const std::string dir = std::string(argv[1]) + "/";
  try {
    kgv::Context ctx(0);
    // --- MuHash: elements.txt -> finalize
    {
      std::ifstream f(dir + "elements.txt");
      std::string line;
      std::vector<std::vector<uint8_t>> add, rem;
      kgv::MuHash empty(ctx);
      std::cout << "muhash_empty " << hex(empty.finalize().data(), 32) << "\n";
      kgv::MuHash m(ctx);
      while (std::getline(f, line)) {
        if (line.empty()) continue;
        if (line[0] == '-') rem.push_back(unhex(line.substr(1))); else add.push_back(unhex(line));
      }
      m.update(add, rem);
      std::cout << "muhash_elements " << hex(m.finalize().data(), 32) << "\n";
    }
    // --- signatures
    {
      auto t = slurp<uint8_t>(dir + "triples.bin");
      size_t n = t.size() / 128;
      std::vector<uint8_t> pk(32 * n), msg(32 * n), sig(64 * n);
      for (size_t i = 0; i < n; i++) {
        std::memcpy(&pk[32 * i], &t[128 * i], 32); std::memcpy(&msg[32 * i], &t[128 * i + 32], 32); std::memcpy(&sig[64 * i], &t[128 * i + 64], 64);
      }
      auto st = kgv::SigVerifier(ctx).check_schnorr_signatures(pk, msg, sig);
      std::cout << "schnorr";
      for (uint8_t s : st) std::cout << " " << (int)s;
      std::cout << "\n";
    }
    // --- transactions: populated validation, then against a UTXO set with MuHash, then the set commitment identity
    kgv::TxBatch b;
    b.assign(slurp<kgv_tx>(dir + "txs.bin"), slurp<kgv_input>(dir + "inputs.bin"), slurp<kgv_output>(dir + "outputs.bin"), slurp<kgv_utxo_entry>(dir + "entries.bin"),
             slurp<uint8_t>(dir + "arena.bin"));
    kgv::Params prm;
    if (argc > 2) prm.storage_mass_parameter = std::stoull(argv[2]);
    kgv::TransactionValidator tv(ctx, prm);
    auto res = tv.validate_populated_transactions(b, 10);
    std::cout << "populated";
    for (auto& r : res) std::cout << " " << (int)r.status << ":" << (int)r.script_err << ":" << r.fee;
    std::cout << "\n";
    kgv::UtxoSet us(ctx, 1 << 14);
    {
      auto keys = slurp<uint8_t>(dir + "fund_keys.bin");
      auto ents = slurp<kgv_utxo_entry>(dir + "fund_entries.bin");
      auto arena = slurp<uint8_t>(dir + "fund_arena.bin");
      std::vector<std::pair<kgv::TransactionOutpoint, kgv::UtxoEntry>> added;
      for (size_t i = 0; i < ents.size(); i++) {
        kgv::TransactionOutpoint o;
        std::memcpy(o.transaction_id.data(), &keys[36 * i], 32);
        o.index = (uint32_t)keys[36 * i + 32] | ((uint32_t)keys[36 * i + 33] << 8) | ((uint32_t)keys[36 * i + 34] << 16) | ((uint32_t)keys[36 * i + 35] << 24);
        kgv::UtxoEntry e;
        e.amount = ents[i].amount; e.block_daa_score = ents[i].block_daa_score; e.is_coinbase = ents[i].is_coinbase != 0;
        e.script_public_key.version = ents[i].spk_version;
        e.script_public_key.script.assign(arena.begin() + ents[i].script_off, arena.begin() + ents[i].script_off + ents[i].script_len);
        added.emplace_back(o, e);
      }
      us.write_diff({}, added);
    }
    std::cout << "utxo_count " << us.count() << "\n";
    kgv::MuHash before = us.muhash();
    auto vm = tv.validate_transactions_with_muhash_in_parallel(us, b, 10);
    std::cout << "in_parallel";
    std::vector<uint8_t> accept;
    for (auto& r : vm.first) { std::cout << " " << (int)r.status; accept.push_back(r.status == KGV_TX_OK); }
    std::cout << "\n";
    std::cout << "tx_muhash_num " << hex(vm.second.numerator().data(), 384) << "\n";
    std::cout << "tx_muhash_den " << hex(vm.second.denominator().data(), 384) << "\n";
    // --- composed view + SigCache: the same block through base ∘ (diff layer), twice with a cache attached (second pass answered from the cache);
    //     applying it to the LAYER leaves the base untouched; discard drops the branch
    {
      kgv::UtxoSet view(ctx, us, 1 << 12);
      kgv::SigCache cache(ctx, 1 << 12);
      auto r1 = tv.validate_transactions_in_parallel(view, b, 10);
      auto k1 = cache.counters();
      auto r2 = tv.validate_transactions_in_parallel(view, b, 10);
      auto k2 = cache.counters();
      bool same = true;
      for (size_t i = 0; i < r1.size(); i++) same = same && r1[i].status == vm.first[i].status && r2[i].status == vm.first[i].status && r1[i].script_err == vm.first[i].script_err;
      view.add_transactions(b, accept, 10);
      auto r3 = tv.validate_transactions_in_parallel(view, b, 10);  // everything accepted is now spent IN THE VIEW
      size_t missing = 0;
      for (size_t i = 0; i < r3.size(); i++) missing += (accept[i] && r3[i].status == KGV_TX_MISSING_OUTPOINTS);
      size_t n_acc = 0;
      for (uint8_t a : accept) n_acc += a;
      std::cout << "view " << (same ? 1 : 0) << " " << us.count() << " " << (k2.get_counts - k1.get_counts) << " " << k1.insert_counts << " " << (k2.insert_counts - k1.insert_counts) << " "
                << missing << " " << n_acc << "\n";
      view.discard();
    }
    us.add_transactions(b, accept, 10);
    before.combine(vm.second);
    std::cout << "commitment_matches " << (before.finalize() == us.muhash().finalize() ? 1 : 0) << "\n";
    // --- pruning-point import through the mirror: the set leaves `us` through its iterator and enters a fresh set chunk by chunk
    //     (append_imported_pruning_point_utxos); the imported multiset must finalize to the source's commitment
    {
      auto all = us.iterator();
      kgv::UtxoSet imported(ctx, 1 << 12);
      kgv::MuHash ms(ctx);
      for (size_t a = 0; a < all.size(); a += 7) {
        std::vector<std::pair<kgv::TransactionOutpoint, kgv::UtxoEntry>> chunk(all.begin() + a, all.begin() + std::min(all.size(), a + 7));
        imported.append_imported_pruning_point_utxos(chunk, ms);
      }
      std::cout << "pruning_import " << all.size() << " " << imported.count() << " " << (ms.finalize() == us.muhash().finalize() ? 1 : 0) << " "
                << (imported.muhash().finalize() == us.muhash().finalize() ? 1 : 0) << "\n";
    }
    // --- block bodies
    auto first = slurp<uint32_t>(dir + "blocks.bin");
    auto roots = kgv::calc_hash_merkle_roots(ctx, b, first);
    std::cout << "merkle";
    for (auto& r : roots) std::cout << " " << hex(r.data(), 32);
    std::cout << "\n";
    auto chk = kgv::check_block_bodies(ctx, b, first);
    std::cout << "bodies";
    for (auto& c : chk) std::cout << " " << c.status << ":" << c.index;
    std::cout << "\n";
    // --- error behaviour: a transport error throws, verdicts never do
    try {
      std::vector<uint8_t> pk(32), msg(64), sig(64);
      kgv::SigVerifier(ctx).check_schnorr_signatures(pk, msg, sig);
      std::cout << "size_mismatch not_detected\n";
    } catch (const kgv::Error& e) {
      std::cout << "size_mismatch throws " << e.code() << "\n";
    }
  } catch (const std::exception& e) {
    std::cout << "EXCEPTION " << e.what() << "\n";
    return 1;
  }
  return 0;
}
