// build/read_mapping_build/bidir_index — drop-in for the executable built from
// VARSCOT_pipeline/read_mapping/bidir_index.cpp (CMakeLists.txt:20-21), same argv (VARSCOT:307).
#include "../../include/varscot_scan.h"
int main(int argc, char **argv) { return vs_bidir_index_main(argc, argv); }
