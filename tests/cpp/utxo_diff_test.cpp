// Runs the reference's UtxoDiff rule table (tests/golden/utxo_diff_rules.json, flattened to text by tests/test_utxo_diff.py)
// through kgv::UtxoDiff of include/kgv.hpp.  No GPU needed.  Input lines: "<this.add> <this.remove> <other.add> <other.remove>"
// with each field a string over {1,2} or "-"; output per line: "<diff_from result> | <with_diff result> | <round trips ok>".
#include <iostream>
#include <sstream>
#include "../../include/kgv.hpp"
using namespace kgv;

static UtxoEntry entry(char c) {
  UtxoEntry e;
  e.amount = c == '1' ? 10 : 20; e.block_daa_score = c == '1' ? 0 : 1; e.is_coinbase = true;
  return e;
}
static UtxoDiff build(const std::string& add, const std::string& rem) {
  UtxoDiff d;
  TransactionOutpoint o;  // (0^32, 0)
  for (char c : add) if (c != '-') d.add[o] = entry(c);
  for (char c : rem) if (c != '-') d.remove[o] = entry(c);
  return d;
}
static std::string show(const UtxoDiff& d) {
  auto f = [](const UtxoCollection& c) { std::string s; for (auto& kv : c) s.push_back(kv.second.amount == 10 ? '1' : '2'); return s.empty() ? std::string("-") : s; };
  return "ok " + f(d.add) + " " + f(d.remove);
}
int main() {
  std::string line;
  while (std::getline(std::cin, line)) {
    std::istringstream is(line);
    std::string ta, tr, oa, orr;
    if (!(is >> ta >> tr >> oa >> orr)) continue;
    UtxoDiff t = build(ta, tr), o = build(oa, orr);
    bool trips = true;
    std::string a, b;
    try { UtxoDiff r = t.diff_from(o); a = show(r); trips = trips && t.with_diff(r) == o; } catch (const UtxoAlgebraError& e) { a = std::string("err ") + e.what(); }
    try { UtxoDiff r = t.with_diff(o); b = show(r); trips = trips && t.diff_from(r) == o; } catch (const UtxoAlgebraError& e) { b = std::string("err ") + e.what(); }
    std::cout << a << " | " << b << " | " << (trips ? 1 : 0) << "\n";
  }
  return 0;
}
