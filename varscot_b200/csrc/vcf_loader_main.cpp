// build/variant_processing_build/vcf_loader — drop-in for the executable built from
// VARSCOT_pipeline/variant_processing/vcf_loader.cpp (called at VARSCOT_pipeline/VARSCOT:274), same argv.
#include "../../include/varscot_scan.h"
int main(int argc, char **argv) { return vs_vcf_loader_main(argc, argv); }
