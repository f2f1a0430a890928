// soda-cr: native computation-reuse scheduler behind the JSON contract the
// reference uses for its external tool (reference:
// src/soda/optimization/computation_reuse.py:1692-1743, ExternalSchedules):
//
//   stdin : {"rattrs": [r0, r1, ...], "aattrs": [a0, a1, ...],
//            "linearizer": {"maxs": [...], "mins": [...], "sizes": [...]},  (optional)
//            "num_pruned": N}                                               (optional)
//   flags : --greedy | --beam | --brute-force   (search effort)
//   stdout: {"rattrs": [...], "left": <int|tree>, "right": <int|tree>,
//            "distance": d, "num_ops": n}
//           a leaf is the operand's aattr tag; "distance" is the offset of the
//           right child's first operand relative to the left child's
//           (reference: make_schedule_from_json, :871-896).
//
// The reference's tool itself (repo Blaok/soda-cr, C++) is not part of the
// reference tree; this is an independent implementation of the same search
// problem: find a binary tree over the operands of a reduction that minimises
// the number of sub-trees that are distinct up to translation.  It is the C++
// twin of soda_b200/optimization/computation_reuse.py (same moves, same
// ordering), which selects it when the binary has been built.
//
// Build: g++ -O2 -std=c++17 -o soda-cr soda_cr.cpp
#include <algorithm>
#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <iostream>
#include <iterator>
#include <map>
#include <memory>
#include <set>
#include <sstream>
#include <string>
#include <vector>

namespace {

// ---- a tiny JSON reader for the fixed input shape ---------------------------------
struct Input {
  std::vector<long long> rattrs;
  std::vector<long long> aattrs;
  std::vector<long long> mins, maxs, sizes;
  int num_pruned = 0;
};

std::vector<long long> ParseArray(const std::string& text, size_t* pos) {
  std::vector<long long> values;
  while (*pos < text.size() && text[*pos] != '[') ++*pos;
  ++*pos;
  std::string token;
  for (; *pos < text.size() && text[*pos] != ']'; ++*pos) {
    char c = text[*pos];
    if (c == ',' || std::isspace(static_cast<unsigned char>(c))) {
      if (!token.empty()) values.push_back(std::atoll(token.c_str()));
      token.clear();
    } else {
      token.push_back(c);
    }
  }
  if (!token.empty()) values.push_back(std::atoll(token.c_str()));
  ++*pos;
  return values;
}

Input ParseInput(const std::string& text) {
  Input in;
  auto find_array = [&](const char* key, std::vector<long long>* out) {
    size_t pos = text.find(std::string("\"") + key + "\"");
    if (pos == std::string::npos) return;
    *out = ParseArray(text, &pos);
  };
  find_array("rattrs", &in.rattrs);
  find_array("aattrs", &in.aattrs);
  find_array("mins", &in.mins);
  find_array("maxs", &in.maxs);
  find_array("sizes", &in.sizes);
  size_t pos = text.find("\"num_pruned\"");
  if (pos != std::string::npos) {
    pos = text.find(':', pos);
    in.num_pruned = std::atoi(text.c_str() + pos + 1);
  }
  if (in.aattrs.empty()) in.aattrs.assign(in.rattrs.size(), 1);
  return in;
}

// ---- patterns: sub-trees up to translation ------------------------------------------
struct Pattern;
using PatternPtr = std::shared_ptr<const Pattern>;
using Leaves = std::vector<std::pair<long long, long long>>;  // sorted (offset, tag)

struct Pattern {
  Leaves leaves;  // relative to the base (offset of the left-most leaf)
  PatternPtr left, right;
  long long distance = 0;
  bool is_leaf() const { return left == nullptr; }
};

struct LeavesLess {
  bool operator()(const Leaves& a, const Leaves& b) const { return a < b; }
};

PatternPtr MakeLeaf(long long tag) {
  auto p = std::make_shared<Pattern>();
  p->leaves = {{0, tag}};
  return p;
}

PatternPtr Merge(const PatternPtr& left, const PatternPtr& right,
                 long long distance) {
  auto p = std::make_shared<Pattern>();
  p->leaves = left->leaves;
  for (const auto& leaf : right->leaves)
    p->leaves.emplace_back(leaf.first + distance, leaf.second);
  std::sort(p->leaves.begin(), p->leaves.end());
  p->left = left;
  p->right = right;
  p->distance = distance;
  return p;
}

void CollectSubtrees(const PatternPtr& p, std::set<Leaves, LeavesLess>* seen) {
  if (p->is_leaf()) return;
  CollectSubtrees(p->left, seen);
  CollectSubtrees(p->right, seen);
  seen->insert(p->leaves);
}

int NumOps(const PatternPtr& p) {
  std::set<Leaves, LeavesLess> seen;
  CollectSubtrees(p, &seen);
  return static_cast<int>(seen.size());
}

struct Item {
  long long base;
  PatternPtr pattern;
};

bool ItemLess(const Item& a, const Item& b) {
  if (a.base != b.base) return a.base < b.base;
  return a.pattern->leaves < b.pattern->leaves;
}

struct Move {
  PatternPtr left, right;
  long long distance;
  std::vector<std::pair<int, int>> occurrences;
  bool aligned = false;
};

std::vector<std::pair<int, int>> DisjointOccurrences(
    const std::vector<Item>& items, const PatternPtr& left,
    const PatternPtr& right, long long distance) {
  std::map<std::pair<long long, Leaves>, std::vector<int>> position;
  for (int i = 0; i < static_cast<int>(items.size()); ++i)
    position[{items[i].base, items[i].pattern->leaves}].push_back(i);
  std::vector<int> order(items.size());
  for (int i = 0; i < static_cast<int>(items.size()); ++i) order[i] = i;
  std::sort(order.begin(), order.end(),
            [&](int a, int b) { return ItemLess(items[a], items[b]); });
  std::vector<char> used(items.size(), 0);
  std::vector<std::pair<int, int>> found;
  for (int i : order) {
    if (used[i] || items[i].pattern->leaves != left->leaves) continue;
    auto it = position.find({items[i].base + distance, right->leaves});
    if (it == position.end()) continue;
    for (int j : it->second) {
      if (j != i && !used[j]) {
        used[i] = used[j] = 1;
        found.emplace_back(i, j);
        break;
      }
    }
  }
  return found;
}

struct Linearizer {
  std::vector<long long> weights;  // empty: 1-D
  // dimension in which two offsets differ if they differ in exactly one, else -1
  int AlignedDim(long long distance) const {
    if (weights.empty()) return 0;
    int dim = -1;
    long long rest = distance;
    for (int d = static_cast<int>(weights.size()) - 1; d >= 0; --d) {
      long long q = rest / weights[d];
      // offsets are differences of non-negative coordinates: round to nearest
      long long r = rest - q * weights[d];
      if (d > 0 && 2 * std::llabs(r) > weights[d]) {
        q += r > 0 ? 1 : -1;
        r = rest - q * weights[d];
      }
      if (q != 0) {
        if (dim != -1) return -1;
        dim = d;
      }
      rest = r;
    }
    return dim;
  }
};

std::vector<Move> CandidateMoves(const std::vector<Item>& items, int limit,
                                 const Linearizer& linearizer) {
  struct Key {
    Leaves left, right;
    long long distance;
    bool operator<(const Key& o) const {
      if (left != o.left) return left < o.left;
      if (right != o.right) return right < o.right;
      return distance < o.distance;
    }
  };
  struct Entry {
    int count = 0;
    PatternPtr left, right;
  };
  std::map<Key, Entry> counts;
  for (size_t i = 0; i < items.size(); ++i) {
    for (size_t j = i + 1; j < items.size(); ++j) {
      const Item* a = &items[i];
      const Item* b = &items[j];
      if (ItemLess(*b, *a)) std::swap(a, b);
      Key key{a->pattern->leaves, b->pattern->leaves, b->base - a->base};
      Entry& entry = counts[key];
      ++entry.count;
      entry.left = a->pattern;
      entry.right = b->pattern;
    }
  }
  std::vector<std::pair<Key, Entry>> ranked(counts.begin(), counts.end());
  std::stable_sort(ranked.begin(), ranked.end(),
                   [](const auto& a, const auto& b) {
                     if (a.second.count != b.second.count)
                       return a.second.count > b.second.count;
                     return std::llabs(a.first.distance) <
                            std::llabs(b.first.distance);
                   });
  std::vector<Move> moves;
  size_t best = 0;
  for (const auto& kv : ranked) {
    if (kv.second.count < 2) break;
    if (static_cast<int>(moves.size()) >= limit &&
        static_cast<size_t>(kv.second.count) < best)
      break;
    if (static_cast<int>(moves.size()) >= 4 * limit) break;
    auto occ = DisjointOccurrences(items, kv.second.left, kv.second.right,
                                   kv.first.distance);
    if (occ.size() >= 2) {
      best = std::max(best, occ.size());
      moves.push_back(
          Move{kv.second.left, kv.second.right, kv.first.distance, occ, false});
    }
  }
  std::stable_sort(moves.begin(), moves.end(), [](const Move& a, const Move& b) {
    if (a.occurrences.size() != b.occurrences.size())
      return a.occurrences.size() > b.occurrences.size();
    return std::llabs(a.distance) < std::llabs(b.distance);
  });
  if (static_cast<int>(moves.size()) > limit) moves.resize(limit);
  // reuse along a single dimension, highest dimension first
  int dims = linearizer.weights.empty() ? 1
                                        : static_cast<int>(linearizer.weights.size());
  for (int axis = dims - 1; axis >= 0; --axis) {
    std::vector<Move> extra;
    int examined = 0;
    for (const auto& kv : ranked) {
      if (kv.second.count < 2 || examined >= 2 * limit) break;
      if (linearizer.AlignedDim(kv.first.distance) != axis) continue;
      ++examined;
      auto occ = DisjointOccurrences(items, kv.second.left, kv.second.right,
                                     kv.first.distance);
      if (occ.size() >= 2)
        extra.push_back(
            Move{kv.second.left, kv.second.right, kv.first.distance, occ, true});
    }
    if (extra.empty()) continue;
    std::stable_sort(extra.begin(), extra.end(),
                     [](const Move& a, const Move& b) {
                       return a.occurrences.size() > b.occurrences.size();
                     });
    for (size_t k = 0; k < extra.size() && k < 2; ++k) {
      bool known = false;
      for (Move& m : moves) {
        if (m.left->leaves == extra[k].left->leaves &&
            m.right->leaves == extra[k].right->leaves &&
            m.distance == extra[k].distance) {
          m.aligned = true;
          known = true;
        }
      }
      if (!known) moves.push_back(extra[k]);
    }
    break;
  }
  return moves;
}

std::vector<Item> Apply(const std::vector<Item>& items, const Move& move) {
  PatternPtr merged = Merge(move.left, move.right, move.distance);
  std::vector<char> consumed(items.size(), 0);
  std::vector<Item> result;
  for (const auto& occ : move.occurrences) {
    consumed[occ.first] = consumed[occ.second] = 1;
    result.push_back(Item{items[occ.first].base, merged});
  }
  for (size_t k = 0; k < items.size(); ++k)
    if (!consumed[k]) result.push_back(items[k]);
  return result;
}

PatternPtr Finish(std::vector<Item> items) {
  std::sort(items.begin(), items.end(), ItemLess);
  PatternPtr tree = items[0].pattern;
  for (size_t k = 1; k < items.size(); ++k)
    tree = Merge(tree, items[k].pattern, items[k].base - items[0].base);
  return tree;
}

struct State {
  std::vector<Item> items;
  bool regular;
  int bound;  // operations so far + items - 1
};

int OpsSoFar(const std::vector<Item>& items) {
  std::set<Leaves, LeavesLess> seen;
  for (const Item& item : items) CollectSubtrees(item.pattern, &seen);
  return static_cast<int>(seen.size());
}

PatternPtr Search(const std::vector<Item>& start, int beam_width, int branch,
                  const Linearizer& linearizer) {
  std::vector<State> frontier{{start, true, 0}};
  std::vector<PatternPtr> finished;
  std::set<std::vector<std::pair<long long, Leaves>>> seen;
  while (!frontier.empty()) {
    std::vector<State> next;
    for (const State& state : frontier) {
      auto moves = CandidateMoves(state.items, branch, linearizer);
      if (moves.empty()) {
        finished.push_back(Finish(state.items));
        continue;
      }
      for (const Move& move : moves) {
        std::vector<Item> items = Apply(state.items, move);
        std::vector<std::pair<long long, Leaves>> key;
        for (const Item& item : items)
          key.emplace_back(item.base, item.pattern->leaves);
        std::sort(key.begin(), key.end());
        if (!seen.insert(key).second) continue;
        int bound = OpsSoFar(items) + static_cast<int>(items.size()) - 1;
        next.push_back(State{std::move(items), state.regular && move.aligned,
                             bound});
      }
    }
    std::stable_sort(next.begin(), next.end(), [](const State& a, const State& b) {
      if (a.bound != b.bound) return a.bound < b.bound;
      return a.items.size() < b.items.size();
    });
    frontier.clear();
    int protected_left = 2;
    for (size_t k = 0; k < next.size(); ++k) {
      if (static_cast<int>(k) < beam_width) {
        frontier.push_back(std::move(next[k]));
      } else if (next[k].regular && protected_left > 0) {
        frontier.push_back(std::move(next[k]));
        --protected_left;
      }
    }
  }
  PatternPtr best;
  int best_ops = 1 << 30;
  for (const PatternPtr& tree : finished) {
    int ops = NumOps(tree);
    if (ops < best_ops) {
      best_ops = ops;
      best = tree;
    }
  }
  return best;
}

// ---- exhaustive search (--optimal) -----------------------------------------------
// Every binary tree over the operands, as the reference's CommSchedules does
// (reference :983-1059: the first operand stays in the left child, every
// subset of the others joins it, both children are expanded recursively), with
// the running count of distinct sub-trees as the pruning cost (the reference's
// "skip-with-partial-cost").  (2n-3)!! trees: exact up to 9 operands in
// seconds; beyond that the node budget ends the search, the best tree so far
// is returned and "proven_optimal" is false (the reference gives up after a
// 300 s timeout in the same way).  All trees with the fewest operations are
// kept (up to a cap) so that the caller can break the tie by total reuse
// distance, the second component of the reference's cost.
struct Exhaustive {
  std::vector<long long> rattrs, aattrs;  // sorted by rattr
  int best_ops = 1 << 30;
  std::vector<PatternPtr> best_trees;
  std::set<std::string> best_texts;
  std::map<Leaves, int, LeavesLess> live;  // distinct sub-trees on the DFS path
  long long nodes = 0, node_limit = 0;
  bool complete = true;
  size_t cap = 4096;

  static std::string Text(const PatternPtr& p) {
    if (p->is_leaf()) return std::to_string(p->leaves[0].second);
    return "(" + Text(p->left) + "=" + std::to_string(p->distance) + "=" +
           Text(p->right) + ")";
  }

  // (std::function, not a template: every nesting level would be a new type)
  using Cont = std::function<void(const PatternPtr&)>;
  void Enumerate(const std::vector<int>& set, const Cont& cont) {
    if (!complete) return;
    if (set.size() == 1) {
      cont(MakeLeaf(aattrs[set[0]]));
      return;
    }
    const int n = static_cast<int>(set.size());
    // subsets of set[1..] that join set[0] on the left; not all of them
    for (unsigned mask = 0; mask + 1 < (1u << (n - 1)); ++mask) {
      std::vector<int> left{set[0]}, right;
      for (int i = 1; i < n; ++i)
        ((mask >> (i - 1)) & 1u ? left : right).push_back(set[i]);
      const long long distance = rattrs[right[0]] - rattrs[left[0]];
      Enumerate(left, [&](const PatternPtr& l) {
        Enumerate(right, [&](const PatternPtr& r) {
          if (++nodes > node_limit) {
            complete = false;
            return;
          }
          PatternPtr merged = Merge(l, r, distance);
          int& count = live[merged->leaves];
          ++count;
          // one more operation is certain unless this is the whole tree
          const int ops = static_cast<int>(live.size());
          const bool whole = merged->leaves.size() == rattrs.size();
          if (ops + (whole ? 0 : 1) <= best_ops) cont(merged);
          if (--live[merged->leaves] == 0) live.erase(merged->leaves);
        });
      });
    }
  }

  void Run() {
    std::vector<int> all(rattrs.size());
    for (size_t i = 0; i < all.size(); ++i) all[i] = static_cast<int>(i);
    Enumerate(all, [&](const PatternPtr& tree) {
      const int ops = static_cast<int>(live.size());
      if (ops < best_ops) {
        best_ops = ops;
        best_trees.clear();
        best_texts.clear();
      }
      if (ops == best_ops && best_trees.size() < cap &&
          best_texts.insert(Text(tree)).second)
        best_trees.push_back(tree);
    });
  }
};

void PrintTree(const PatternPtr& p, std::ostream& out) {
  if (p->is_leaf()) {
    out << p->leaves[0].second;
    return;
  }
  out << "{\"left\": ";
  PrintTree(p->left, out);
  out << ", \"right\": ";
  PrintTree(p->right, out);
  out << ", \"distance\": " << p->distance << "}";
}

}  // namespace

int main(int argc, char** argv) {
  int beam_width = 6, branch = 4;
  bool optimal = false;
  for (int i = 1; i < argc; ++i) {
    std::string flag = argv[i];
    if (flag == "--greedy") {
      beam_width = 1;
      branch = 1;
    } else if (flag == "--beam") {
      beam_width = 6;
      branch = 4;
    } else if (flag == "--brute-force") {
      beam_width = 256;
      branch = 32;
    } else if (flag == "--optimal") {
      optimal = true;
      beam_width = 256;  // the beam result seeds the pruning bound
      branch = 32;
    } else if (flag == "--help") {
      std::cout << "usage: soda-cr [--greedy|--beam|--brute-force|--optimal] < in.json\n";
      return 0;
    }
  }
  std::string text((std::istreambuf_iterator<char>(std::cin)),
                   std::istreambuf_iterator<char>());
  Input in = ParseInput(text);
  if (in.rattrs.size() < 2 || in.rattrs.size() != in.aattrs.size()) {
    std::cerr << "soda-cr: need at least two rattrs and as many aattrs\n";
    return 1;
  }
  Linearizer linearizer;
  if (!in.sizes.empty()) {
    long long weight = 1;
    for (size_t d = 0; d < in.sizes.size(); ++d) {
      linearizer.weights.push_back(weight);
      weight *= in.sizes[d];
    }
  }
  const size_t n = in.rattrs.size();
  if (n <= 7) {
    beam_width = std::max(beam_width, 64);
    branch = std::max(branch, 16);
  } else if (n > 128) {
    beam_width = 1;
    branch = 1;
  } else if (n > 48) {
    beam_width = std::min(beam_width, 2);
    branch = std::min(branch, 2);
  }
  std::vector<Item> start;
  for (size_t i = 0; i < n; ++i)
    start.push_back(Item{in.rattrs[i], MakeLeaf(in.aattrs[i])});
  PatternPtr best = Search(start, beam_width, branch, linearizer);
  std::vector<PatternPtr> alternatives;
  bool proven = false;
  long long nodes = 0;
  if (optimal) {
    Exhaustive search;
    std::vector<size_t> order(n);
    for (size_t i = 0; i < n; ++i) order[i] = i;
    std::sort(order.begin(), order.end(), [&](size_t a, size_t b) {
      return in.rattrs[a] < in.rattrs[b];
    });
    for (size_t i : order) {
      search.rattrs.push_back(in.rattrs[i]);
      search.aattrs.push_back(in.aattrs[i]);
    }
    search.best_ops = NumOps(best);
    const char* limit = getenv("SODA_CR_NODE_LIMIT");
    search.node_limit = limit ? atoll(limit) : 40000000LL;
    search.Run();
    proven = search.complete;
    nodes = search.nodes;
    if (!search.best_trees.empty()) {
      best = search.best_trees[0];
      alternatives.assign(search.best_trees.begin() + 1, search.best_trees.end());
    }
    if (!proven)
      std::cerr << "soda-cr: --optimal stopped after " << nodes
                << " trees; the schedule is the best found, not proven optimal\n";
  }
  std::vector<long long> sorted_rattrs = in.rattrs;
  std::sort(sorted_rattrs.begin(), sorted_rattrs.end());
  std::cout << "{\"rattrs\": [";
  for (size_t i = 0; i < sorted_rattrs.size(); ++i)
    std::cout << (i ? ", " : "") << sorted_rattrs[i];
  std::cout << "], \"num_ops\": " << NumOps(best) << ", \"left\": ";
  PrintTree(best->left, std::cout);
  std::cout << ", \"right\": ";
  PrintTree(best->right, std::cout);
  std::cout << ", \"distance\": " << best->distance;
  if (optimal) {
    std::cout << ", \"proven_optimal\": " << (proven ? "true" : "false")
              << ", \"trees\": " << nodes << ", \"alternatives\": [";
    for (size_t i = 0; i < alternatives.size(); ++i) {
      std::cout << (i ? ", " : "");
      PrintTree(alternatives[i], std::cout);
    }
    std::cout << "]";
  }
  std::cout << "}\n";
  return 0;
}
