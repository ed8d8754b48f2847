# builds tuning variants of libvarscot_scan.so into build/variants/ (here, without a GPU; they travel with gpurun)
# usage: bash tools/build_variants.sh name1 "flags1" name2 "flags2" ...
cd "$(dirname "$0")/.."
mkdir -p build/variants
NVCC=/usr/local/cuda/bin/nvcc
SRCS="varscot_b200/csrc/vs_device.cu varscot_b200/csrc/vs_host.cpp varscot_b200/csrc/vs_cli.cpp varscot_b200/csrc/vs_vcf.cpp varscot_b200/csrc/vs_merge.cpp"
while [ $# -ge 2 ]; do
  name="$1"; flags="$2"; shift 2
  ( $NVCC -gencode arch=compute_100a,code=sm_100a $flags -O3 -std=c++17 -lineinfo -ccbin /usr/bin/g++ -Xcompiler -fPIC,-Wall,-Wno-unused-function --cudart static \
      -Xptxas -v -shared -o build/variants/lib_$name.so $SRCS -lpthread 2>&1 | grep -A1 "k_scoreILi6" | grep -E "Used|spill" | tr '\n' ' '; echo " <- $name" ) &
done
wait
