#!/bin/sh
# tools/build_variant.sh NAME "EXTRA NVCC FLAGS" [ALTERNATIVE SOURCE] [UNIT] - A/B builds of one translation unit
# (UNIT, default demod_kernels; e.g. demod_fast, decode_kernels, fcch_kernels): compiles csrc/UNIT.cu (or the
# alternative source) with extra flags and links osmo_gmr_b200/build/variants/libNAME.so from it plus the regular
# objects.  Select at run time with GMR1B200_LIB=<path> (developer knob of lib.py).
set -e
cd "$(dirname "$0")/../osmo_gmr_b200"
make -j8 >/dev/null
mkdir -p build/variants
nvcc -O3 -std=c++17 --expt-relaxed-constexpr -gencode arch=compute_100a,code=sm_100a -lineinfo \
     -Xcompiler -fPIC $2 -Icsrc -c ${3:-csrc/${4:-demod_kernels}.cu} -o build/variants/unit_$1.o
OBJS=$(ls build/*.o | grep -v "/${4:-demod_kernels}.o")
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o build/variants/lib$1.so $OBJS build/variants/unit_$1.o \
     -lcudart -Xlinker -Bsymbolic-functions
echo build/variants/lib$1.so
