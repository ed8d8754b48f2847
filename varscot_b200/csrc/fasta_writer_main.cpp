// build/variant_processing_build/fasta_writer — drop-in for the executable built from
// VARSCOT_pipeline/variant_processing/fasta_writer.cpp (called at VARSCOT_pipeline/VARSCOT:260), same argv.
#include "../../include/varscot_scan.h"
int main(int argc, char **argv) { return vs_fasta_writer_main(argc, argv); }
