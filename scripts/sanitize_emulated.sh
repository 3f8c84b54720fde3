#!/bin/bash
# The parity suite on the emulated kernels under AddressSanitizer, with exact-size "device" buffers filled with garbage:
# out-of-bounds accesses and reliance on zero-initialised memory in the kernels show up here (test infrastructure, CPU only).
set -e
cd "$(dirname "$0")/.."
make -C tests/cuda_emu libtfbs_emu_asan.so > /dev/null
LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0:detect_stack_use_after_return=0 \
TFBS_B200_LIB=$PWD/tests/cuda_emu/libtfbs_emu_asan.so \
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -p no:cacheprovider -k "${1:-not config2_slice and not config3_like}"
