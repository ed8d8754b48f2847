// build/variant_processing_build/bam_merger_ref_only — drop-in for the executable built from
// VARSCOT_pipeline/variant_processing/bam_merger_ref_only.cpp (called at VARSCOT_pipeline/VARSCOT:338-343), same argv.
#include "../../include/varscot_scan.h"
int main(int argc, char **argv) { return vs_bam_merger_ref_only_main(argc, argv); }
