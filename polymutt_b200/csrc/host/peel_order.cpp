// Elston–Stewart peeling order for one family, the behaviour of ES_Peeling in the reference
// (src/FamilyLikelihoodES.cpp:46-277): work lists of leaves, peripheral spouses and "roof" couples
// are drained in that priority until famSize-1 people are peeled.  Runs once per pedigree on the
// host; the device only sees the resulting (type, from, to) triples.
#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <deque>
#include <map>
#include <utility>
#include <vector>

#include "host_error.h"
#include "polymutt_b200.h"

namespace {

struct Net {
  int n;
  std::vector<std::vector<int>> parents, kids, mates;
  bool is_final(int i) const { return parents[i].empty() && mates[i].empty() && kids[i].empty(); }
  bool is_leaf(int i) const { return kids[i].empty() && mates[i].empty(); }
  bool is_peripheral(int i) const { return kids[i].empty() && parents[i].empty() && mates[i].size() == 1; }
  bool is_roof(int i) const {
    if (mates[i].size() != 1) return false;
    int s = mates[i][0];
    return mates[s].size() == 1 && parents[i].empty() && parents[s].empty() && kids[i].size() == 1 && kids[s].size() == 1;
  }
};

void drop(std::vector<int> &v, int x) {
  auto it = std::find(v.begin(), v.end(), x);
  if (it != v.end()) v.erase(it);
}

using Couple = std::pair<int, int>;
int find_couple(const std::vector<Couple> &v, int a, int b) {  // either orientation, ES:472-483
  for (size_t i = 0; i < v.size(); i++)
    if ((v[i].first == a && v[i].second == b) || (v[i].first == b && v[i].second == a)) return (int)i;
  return -1;
}

}  // namespace

extern "C" int pm_build_peel_order(int32_t n, const int32_t *father, const int32_t *mother, const uint8_t *sex,
                                   pm_peel_step *steps) {
  if (n <= 0 || !father || !mother || !sex || !steps) return pmh::fail(PM_EINVAL, "pm_build_peel_order: bad argument");
  Net g;
  g.n = n;
  g.parents.assign(n, {}); g.kids.assign(n, {}); g.mates.assign(n, {});
  {
    std::map<Couple, int> seen;  // SetupConnections, ES:46-78
    for (int i = 0; i < n; i++) {
      if (father[i] < 0 || mother[i] < 0) continue;
      int fa = father[i], mo = mother[i];
      if (fa >= n || mo >= n) return pmh::fail(PM_EINVAL, "pm_build_peel_order: parent index out of range");
      g.parents[i] = {fa, mo};
      g.kids[fa].push_back(i);
      g.kids[mo].push_back(i);
      if (seen[{fa, mo}]++ == 0) { g.mates[fa].push_back(mo); g.mates[mo].push_back(fa); }
    }
  }
  std::deque<int> leaves, peripherals;
  std::vector<Couple> roofs;
  {
    std::vector<char> in_roof(n, 0);  // BuildInitialPeelable, ES:80-115
    for (int i = 0; i < n; i++) {
      if (g.is_leaf(i)) { leaves.push_back(i); continue; }
      if (g.is_roof(i)) {
        int s = g.mates[i][0];
        if (in_roof[i] || in_roof[s]) continue;
        roofs.push_back(sex[i] == 1 ? Couple(i, s) : Couple(s, i));  // male first
        in_roof[i] = in_roof[s] = 1;
        continue;
      }
      if (g.is_peripheral(i)) peripherals.push_back(i);
    }
  }
  auto add_roof = [&](int who) {  // UpdateRoof, ES:461-470: stored as (who, mate), not male-first
    int s = g.mates[who][0];
    if (find_couple(roofs, who, s) < 0) roofs.emplace_back(who, s);
  };
  int ns = 0, peeled = 0;
  bool done = false;
  auto emit = [&](int type, int f0, int f1, int t0, int t1) {
    steps[ns].type = type; steps[ns].from0 = f0; steps[ns].from1 = f1; steps[ns].to0 = t0; steps[ns].to1 = t1;
    ns++;
  };
  while (!done && !(leaves.empty() && roofs.empty() && peripherals.empty())) {  // BuildPeelingOrder, ES:135-277
    while (!leaves.empty()) {
      int leaf = leaves.front(); leaves.pop_front();
      if (g.parents[leaf].size() != 2) return pmh::fail(PM_EINVAL, "Peeling error for person %d! Check pedigree structure!!", leaf);
      int fa = g.parents[leaf][0], mo = g.parents[leaf][1];
      peeled++;
      emit(PM_PEEL_CHILD_TO_PARENTS, leaf, -1, fa, mo);
      drop(g.kids[fa], leaf);
      drop(g.kids[mo], leaf);
      g.parents[leaf].clear();
      if (g.is_peripheral(fa)) peripherals.push_back(fa);
      if (g.is_peripheral(mo)) peripherals.push_back(mo);
      int pos = find_couple(roofs, fa, mo);
      if (pos > 0) roofs.erase(roofs.begin() + pos);  // a couple sitting at position 0 stays (ES:185-187)
      if (peeled == n - 1) done = true;
    }
    if (done) break;
    while (!peripherals.empty()) {
      int who = peripherals.front(); peripherals.pop_front();
      if (g.mates[who].size() > 1) return pmh::fail(PM_EINVAL, "Peripheral parent can not have more than one spouses!");
      if (g.mates[who].empty()) return pmh::fail(PM_EINVAL, "No spouse can be found for person %d!", who);
      int mate = g.mates[who][0];
      peeled++;
      emit(PM_PEEL_SPOUSE_TO_SPOUSE, who, -1, mate, -1);
      drop(g.mates[mate], who);
      g.mates[who].clear();
      if (g.is_final(mate)) {
        if (peeled != n - 1)
          return pmh::fail(PM_EINVAL, "Are there disconnected sub-pedigrees in the family? Please move sub-pedigrees to separate families.");
        done = true;
        break;
      }
      if (g.is_leaf(mate)) leaves.push_back(mate);
      else if (g.is_peripheral(mate)) peripherals.push_back(mate);
      else if (g.is_roof(mate)) add_roof(mate);
    }
    if (done) break;
    if (!leaves.empty() || !peripherals.empty()) continue;
    while (!roofs.empty()) {
      Couple c = roofs.front();
      roofs.erase(roofs.begin());
      if (g.kids[c.first].size() != 1 || g.kids[c.second].size() != 1)
        return pmh::fail(PM_EINVAL, "Roof can only have one offspring for peeling!");
      int child = g.kids[c.first][0];
      peeled += 2;
      emit(PM_PEEL_PARENTS_TO_CHILD, c.first, c.second, child, -1);
      g.parents[child].clear();
      g.kids[c.first].clear();
      g.kids[c.second].clear();
      if (g.is_peripheral(child)) peripherals.push_back(child);
      else if (g.is_roof(child)) add_roof(child);
      else if (g.is_final(child)) { done = true; break; }
    }
  }
  if (peeled < n - 1) return pmh::fail(PM_EINVAL, "Are there inbreeding loops in the pedigree? It cannot handel inbreeding yet!");
  return ns;
}
