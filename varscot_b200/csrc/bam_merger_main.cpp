// build/variant_processing_build/bam_merger — drop-in for the executable built from
// VARSCOT_pipeline/variant_processing/bam_merger.cpp (called at VARSCOT_pipeline/VARSCOT:332-337), same argv.
#include "../../include/varscot_scan.h"
int main(int argc, char **argv) { return vs_bam_merger_main(argc, argv); }
