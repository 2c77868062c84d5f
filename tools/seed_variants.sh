#!/bin/bash
# Builds libskb variants that differ only in compile-time switches of seed_kernels.cu into build/variants/ (git-ignored,
# travels to the GPU box) and, on a GPU box, times each with tools/seed_bench.py:
#     tools/seed_variants.sh build            # here (no GPU): compile
#     tools/seed_variants.sh run              # on the box: time every variant found
# Variants: SKB_MASK_FMA (1: high-word comparison and hit mask on the FMA pipe through the carry of IMAD.HI), SKB_XS_FMA (0:
# shifts on the ALU pipe; 1: all xor-shift shifts as multiplies on the FMA pipe; 2: only the high-word shifts; 3 / 4: those
# of the seed / marker hash only), SKB_MUL_SPLIT (1: 64-bit multiplies as IMAD + IMAD.HI instead of IMAD.WIDE; 2 / 3: in the seed /
# marker hash only), SKB_SEED_MINBLOCKS (resident CTAs per SM the register allocation is bounded for).
set -e
cd "$(dirname "$0")/.."
V=build/variants
NV="nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -ccbin g++"
if [ "$1" = build ]; then
    mkdir -p $V
    for spec in "mf0_ms0:-DSKB_MASK_FMA=0" "mf1_ms0:-DSKB_MASK_FMA=1" "mf1_ms1:-DSKB_MASK_FMA=1 -DSKB_MUL_SPLIT=1" "mf1_ms2:-DSKB_MASK_FMA=1 -DSKB_MUL_SPLIT=2" \
                "mf1_ms3:-DSKB_MASK_FMA=1 -DSKB_MUL_SPLIT=3" "mf0_ms1:-DSKB_MASK_FMA=0 -DSKB_MUL_SPLIT=1" "mf0_ms3:-DSKB_MASK_FMA=0 -DSKB_MUL_SPLIT=3" $EXTRA_SPECS; do
        name=${spec%%:*}; flags=${spec#*:}
        $NV $flags -c pyskani_b200/csrc/seed_kernels.cu -o $V/seed_$name.o
        $NV -shared -o $V/libskb_$name.so $V/seed_$name.o pyskani_b200/csrc/index_kernels.o pyskani_b200/csrc/screen_kernels.o \
            pyskani_b200/csrc/chain_kernels.o pyskani_b200/csrc/skb_api.o pyskani_b200/csrc/host_pack.o -lcudart -lpthread
        python tools/sass_mix.py $V/seed_$name.o seed_scan_kernel 16 | grep -E "^(alu|fma|total|# issue)" | sed "s/^/$name  /"
    done
else
    for so in $V/libskb_*.so; do
        echo "== $so"
        SKB_LIB=$so python tools/seed_bench.py 100 5000000 | tail -1
    done
fi
