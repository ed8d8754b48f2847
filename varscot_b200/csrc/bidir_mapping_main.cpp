// build/read_mapping_build/bidir_mapping — drop-in for the executable built from
// VARSCOT_pipeline/read_mapping/bidir_mapping.cpp (CMakeLists.txt:23-24), same argv (VARSCOT:296-314).
#include "../../include/varscot_scan.h"
int main(int argc, char **argv) { return vs_bidir_mapping_main(argc, argv); }
