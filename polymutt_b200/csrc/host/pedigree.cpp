#include "pedigree.h"

#include <algorithm>
#include <cctype>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <stdexcept>

namespace pmh {

int natural_compare(const std::string &a, const std::string &b) {
  const char *s = a.c_str(), *t = b.c_str();
  size_t len = a.size();
  for (size_t i = 0; i <= len; i++) {
    int ca = toupper((unsigned char)s[i]), cb = toupper((unsigned char)t[i]);
    if (ca - cb) {
      size_t d = i;
      while (isdigit((unsigned char)s[d]) && isdigit((unsigned char)t[d])) d++;
      if (isdigit((unsigned char)s[d])) return 1;
      if (isdigit((unsigned char)t[d])) return -1;
      return ca - cb;
    }
    if (s[i] == 0) break;
  }
  return 0;
}

static std::vector<std::string> tokenize(const std::string &line, const char *seps) {
  std::vector<std::string> out;
  size_t i = 0, n = line.size();
  while (i < n) {
    while (i < n && strchr(seps, line[i])) i++;
    if (i >= n) break;
    size_t j = i;
    while (j < n && !strchr(seps, line[j])) j++;
    out.emplace_back(line.substr(i, j - i));
    i = j;
  }
  return out;
}

static int translate_sex(const std::string &code) {  // core/PedigreeLoader.cpp:574-599
  switch (code[0]) {
    case 'x': case 'X': case '?': return 0;
    case '1': case 'm': case 'M': return 1;
    case '2': case 'f': case 'F': return 2;
    default: {
      bool result = atoi(code.c_str());  // sic: the reference stores atoi() in a bool
      return result;
    }
  }
}

static std::string slurp(const std::string &path, const char *what) {
  std::ifstream f(path, std::ios::binary);
  if (!f) throw std::runtime_error(std::string(what) + " open for input failed!");
  std::stringstream ss;
  ss << f.rdbuf();
  return ss.str();
}

void Pedigree::load(const std::string &dat_path, const std::string &ped_path) {
  load_from_text(slurp(dat_path, "datFile"), slurp(ped_path, "pedFile"));
}

void Pedigree::load_from_text(const std::string &dat_text, const std::string &ped_text) {
  // .dat: one "<type> <name>" row per data column.  Width in .ped tokens per type
  // (core/PedigreeDescription.cpp:75-137): M = 2 ("/" is a separator), everything else = 1,
  // E = end of file.
  struct Col { char type; std::string name; };
  std::vector<Col> cols;
  {
    std::istringstream in(dat_text);
    std::string line;
    while (std::getline(in, line)) {
      auto tok = tokenize(line, " \t\n\r\f/");
      if (tok.empty()) continue;
      char t = (char)toupper((unsigned char)tok[0][0]);
      if (t == 'E') break;
      cols.push_back({t, tok.size() > 1 ? tok[1] : std::string()});
    }
  }
  int text_cols = 5;
  for (auto &c : cols) text_cols += (c.type == 'M') ? 2 : 1;

  persons.clear();
  families.clear();
  std::istringstream in(ped_text);
  std::string line;
  int lineno = 0;
  while (std::getline(in, line)) {
    auto tok = tokenize(line, " \t\n\r\f/");
    if (tok.empty()) continue;
    if (natural_compare(tok[0], "end") == 0) break;
    lineno++;
    if ((int)tok.size() < text_cols) {
      std::ostringstream msg;
      msg << "Loading Pedigree...\n\nExpecting " << text_cols << " columns,\nbut read only " << tok.size()
          << " columns in line " << lineno << ".";
      throw std::runtime_error(msg.str());
    }
    Person p;
    p.famid = tok[0]; p.pid = tok[1]; p.fatid = tok[2]; p.motid = tok[3];
    p.sex = translate_sex(tok[4]);
    int field = 5;
    bool have_glf = false;
    for (auto &c : cols) {
      if (c.type == 'M') { field += 2; continue; }
      const std::string &v = tok[field++];
      if ((c.type == 'T') && !have_glf && natural_compare(c.name, "GLF_Index") == 0) {
        char *end = nullptr;
        double d = strtod(v.c_str(), &end);
        p.glf_index = (end && *end) ? 0 : (int)d;
        have_glf = true;
      }
    }
    persons.push_back(std::move(p));
  }
  finish();
}

void Pedigree::finish() {
  // Pedigree::Sort, core/Pedigree.cpp:39-85
  std::sort(persons.begin(), persons.end(), [](const Person &a, const Person &b) {
    int r = natural_compare(a.famid, b.famid);
    if (r != 0) return r < 0;
    return natural_compare(a.pid, b.pid) < 0;
  });
  for (size_t i = 1; i < persons.size(); i++)
    if (natural_compare(persons[i - 1].famid, persons[i].famid) == 0 &&
        natural_compare(persons[i - 1].pid, persons[i].pid) == 0)
      throw std::runtime_error("Family " + persons[i].famid + ": Person " + persons[i].pid + " is duplicated");
  auto find = [&](const std::string &famid, const std::string &pid) -> int {
    for (size_t i = 0; i < persons.size(); i++)
      if (natural_compare(persons[i].famid, famid) == 0 && natural_compare(persons[i].pid, pid) == 0) return (int)i;
    return -1;
  };
  // families = runs of equal famid (MakeFamilies, core/Pedigree.cpp:120-142)
  for (int first = 0; first < (int)persons.size();) {
    int last = first;
    while (last < (int)persons.size() && natural_compare(persons[first].famid, persons[last].famid) == 0) last++;
    Family f;
    f.famid = persons[first].famid; f.first = first; f.last = last - 1;
    families.push_back(f);
    first = last;
  }
  // parents are looked up inside the family (FindPerson by famid+pid); a run-local search keeps
  // this linear for thousands of trios
  for (auto &f : families)
    for (int i = f.first; i <= f.last; i++) {
      Person &p = persons[i];
      auto local = [&](const std::string &pid) -> int {
        for (int j = f.first; j <= f.last; j++) if (natural_compare(persons[j].pid, pid) == 0) return j;
        return -1;
      };
      p.father = local(p.fatid);
      p.mother = local(p.motid);
      // CheckParents, core/PedigreePerson.cpp:90-126
      bool both = p.father >= 0 && p.mother >= 0;
      if (!both) {
        if (p.father >= 0 || p.mother >= 0)
          throw std::runtime_error("Parent named " + (p.father < 0 ? p.fatid : p.motid) + " for Person " + p.pid +
                                   " in Family " + p.famid + " is missing");
        p.father = p.mother = -1;
        continue;
      }
      if (persons[p.father].sex == 2 || persons[p.mother].sex == 1) {
        std::swap(p.father, p.mother);
        std::swap(p.fatid, p.motid);
      }
      if (persons[p.father].sex == 2 || persons[p.mother].sex == 1)
        throw std::runtime_error("Parental sex codes don't make sense for Person " + p.pid + " in Family " + p.famid);
    }
  (void)find;
  // Family::Family, core/PedigreeFamily.cpp:11-85
  for (auto &f : families) {
    int count = f.last - f.first + 1;
    f.path.assign(count, -1);
    f.founders = 0;
    for (int i = f.first; i <= f.last; i++)
      if (persons[i].founder()) { persons[i].traverse = f.founders; f.path[f.founders++] = i; }
      else persons[i].traverse = -1;
    f.generations = (count - f.founders) == 0 ? 1 : 2;
    int next = f.founders;
    while (next < count) {
      bool check = false;
      for (int i = f.first; i <= f.last; i++)
        if (persons[i].traverse == -1) {
          int ft = persons[persons[i].father].traverse, mt = persons[persons[i].mother].traverse;
          if (ft >= 0 && mt >= 0) {
            check = true;
            persons[i].traverse = next;
            f.path[next++] = i;
            if (ft >= f.founders || mt >= f.founders) f.generations = 3;
          }
        }
      if (!check)
        throw std::runtime_error("The structure of family " + f.famid + " requires an individual to be his own ancestor.");
    }
  }
  // flat view in VCF column order
  columns_.clear();
  fam_size_.clear(); fam_founders_.clear(); fam_gen_.clear(); father_.clear(); mother_.clear();
  sex_.clear(); peel_first_.clear(); peel_.clear();
  for (auto &f : families) {
    int n = (int)f.path.size();
    fam_size_.push_back(n); fam_founders_.push_back(f.founders); fam_gen_.push_back(f.generations);
    size_t base = father_.size();
    for (int j = 0; j < n; j++) {
      const Person &p = persons[f.path[j]];
      columns_.push_back(f.path[j]);
      sex_.push_back((uint8_t)p.sex);
      father_.push_back(p.father >= 0 ? persons[p.father].traverse : -1);
      mother_.push_back(p.mother >= 0 ? persons[p.mother].traverse : -1);
    }
    peel_first_.push_back((int32_t)peel_.size());
    // every family with non-founders gets an order: extended families always use it, nuclear ones only in VCF
    // mode with a single family (FamilyLikelihoodSeq_VCF.cpp:98-103)
    if (n != f.founders) {
      std::vector<pm_peel_step> steps(n);
      int ns = pm_build_peel_order(n, father_.data() + base, mother_.data() + base, sex_.data() + base, steps.data());
      if (ns < 0) throw std::runtime_error(std::string("family ") + f.famid + ": " + pm_last_error());
      peel_.insert(peel_.end(), steps.begin(), steps.begin() + ns);
    }
  }
  peel_first_.push_back((int32_t)peel_.size());
  view_.n_fam = (int32_t)families.size();
  view_.n_person = (int32_t)persons.size();
  view_.fam_size = fam_size_.data(); view_.fam_founders = fam_founders_.data();
  view_.fam_generations = fam_gen_.data(); view_.sex = sex_.data();
  view_.father = father_.data(); view_.mother = mother_.data();
  view_.peel_first = peel_first_.data(); view_.peel = peel_.empty() ? nullptr : peel_.data();
}

int Pedigree::total_founders() const {
  int n = 0;
  for (auto &f : families) n += f.founders;
  return n;
}

}  // namespace pmh
