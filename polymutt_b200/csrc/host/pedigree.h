// Pedigree model for the host front end: .dat/.ped parsing, the reference's person/family ordering
// rules and the pm_pedigree view handed to the engine.
//
// Behaviour follows core/PedigreeLoader.cpp:14-278 (.ped rows), core/PedigreeDescription.cpp:32-159
// (.dat rows; polymutt only needs "T GLF_Index"), core/Pedigree.cpp:39-85,120-142 (sort + families),
// core/PedigreeFamily.cpp:11-85 (founders-first traversal) and core/PedigreePerson.cpp:90-126
// (swapping mis-sexed parents).  Markers, twins, Mendel/linkage formats are out of scope.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "polymutt_b200.h"

namespace pmh {

struct Person {
  std::string famid, pid, fatid, motid;
  int sex = 0;
  int father = -1, mother = -1;  // indices into Pedigree::persons (sorted order), -1 = founder
  int traverse = -1;             // position inside the family (Family::path order)
  int glf_index = 0;             // value of the GLF_Index trait (0 = no GLF)
  bool founder() const { return father < 0 || mother < 0; }
};

struct Family {
  std::string famid;
  int first = 0, last = 0;       // range in the sorted persons array
  int founders = 0, generations = 1;
  std::vector<int> path;         // persons[] indices in traversal order (founders first)
  bool nuclear() const { return generations == 2 && founders == 2; }
};

class Pedigree {
 public:
  std::vector<Person> persons;   // sorted by natural, case-insensitive (famid, pid)
  std::vector<Family> families;

  // Throws std::runtime_error with the reference's wording on structural problems.
  void load(const std::string &dat_path, const std::string &ped_path);
  void load_from_text(const std::string &dat_text, const std::string &ped_text);

  int n_person() const { return (int)persons.size(); }
  // person index (sorted array) for VCF column c (families in order, members in path order)
  const std::vector<int> &columns() const { return columns_; }

  // Flat arrays backing a pm_pedigree; valid while this object lives.
  const pm_pedigree *view() const { return &view_; }
  int total_founders() const;

 private:
  void finish();
  std::vector<int> columns_;
  std::vector<int32_t> fam_size_, fam_founders_, fam_gen_, father_, mother_, peel_first_;
  std::vector<uint8_t> sex_;
  std::vector<pm_peel_step> peel_;
  pm_pedigree view_{};
};

// String::SlowCompare with NATURAL_ORDERING (core/StringBasics.cpp:431-448)
int natural_compare(const std::string &a, const std::string &b);

}  // namespace pmh
