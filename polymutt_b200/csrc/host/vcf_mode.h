// VCF-input mode of the drop-in executable (--in_vcf): the record loop of PedVCF::VarCallFromVCF
// (src/PedVCF.cpp:42-163), FamilyLikelihoodSeq_VCF::FillPenetrance (src/FamilyLikelihoodSeq_VCF.cpp:267-383:
// bi-allelic check, PL/GL -> capped phred indices) and ::OutputVCF (:437-521), around a batched engine call.
// The LINE_MODE tokenising rules of libVcf (VCFRecord.h:26-131, VCFIndividual.h:26-54, VCFValue.h:150-152) are
// reproduced: tab-separated columns, ':'-separated sample fields, a field is "missing" only when it is
// absent or empty.
#pragma once
#include "driver.h"
#include "params.h"
#include "pedigree.h"

namespace pmh {
int run_vcf_mode(const Options &opt, const Pedigree &ped, const Engine &engine);
}
