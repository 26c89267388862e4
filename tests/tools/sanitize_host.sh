#!/bin/bash
# Host runtime under the sanitizers: builds libpmcts_b200 twice with instrumented host code (address + undefined,
# then thread) and runs the CPU suites that drive the native search (thread pool, trees, master, update blocks,
# generator emulation) against each build; then the C oracle (the parity checker) the same way against its own suites.
# Device code is compiled as usual and not exercised: no GPU needed.
#   tests/tools/sanitize_host.sh            -> prints the reports that name this library (none expected)
set -u
cd "$(dirname "$0")/../.."
P=iboat_pmcts_b200/csrc
GCC_LIB=$(dirname "$(gcc -print-file-name=libasan.so)")
SRC="$P/pmcts_b200.cu $P/pmcts_forest_host.cpp $P/pmcts_search.cu $P/pmcts_master.cu $P/pmcts_peer.cu"
TESTS="tests/test_host_forest.py tests/test_host_tree.py tests/test_abi.py tests/test_forest_gloo.py"
OUT=${TMPDIR:-/tmp}/pmcts_sanitize
rm -rf "$OUT"; mkdir -p "$OUT"
build() {  # name, flags, runtime libs
    nvcc -gencode arch=compute_100a,code=sm_100a -O1 -g -std=c++17 -shared \
        -Xcompiler -fPIC,-fno-omit-frame-pointer,"$2" -o "$OUT/libpmcts_$1.so" $SRC -ldl $3 > "$OUT/build_$1.log" 2>&1 \
        || { echo "build of the $1 library failed: $OUT/build_$1.log"; exit 1; }
}
build asan -fsanitize=address,-fsanitize=undefined "-lasan -lubsan"
PMCTS_B200_LIB=$OUT/libpmcts_asan.so LD_PRELOAD="$GCC_LIB/libasan.so $GCC_LIB/libubsan.so" \
    ASAN_OPTIONS=detect_leaks=0:halt_on_error=0:log_path=$OUT/asan UBSAN_OPTIONS=print_stacktrace=1:log_path=$OUT/ubsan \
    python -m pytest $TESTS -q -m "not gpu" -p no:cacheprovider | tail -1
build tsan -fsanitize=thread -ltsan
# (a process with reports exits 66, which fails tests that start interpreters of their own: the logs are the result)
PMCTS_B200_LIB=$OUT/libpmcts_tsan.so LD_PRELOAD="$GCC_LIB/libtsan.so" PMCTS_HOST_THREADS=8 \
    TSAN_OPTIONS=log_path=$OUT/tsan:halt_on_error=0:report_signal_unsafe=0:exitcode=0 \
    python -m pytest $TESTS -q -m "not gpu" -p no:cacheprovider | tail -1
# the checker itself: the C oracle under address + undefined, against its fixtures, the live reference and the host model
gcc -O1 -g -ffp-contract=off -fno-fast-math -fno-builtin-pow -pthread -fPIC -fsanitize=address,undefined \
    -fno-omit-frame-pointer -shared -o "$OUT/liboracle_asan.so" oracle/pmcts_oracle.c -lm \
    || { echo "build of the oracle failed"; exit 1; }
PMCTS_ORACLE_LIB=$OUT/liboracle_asan.so LD_PRELOAD="$GCC_LIB/libasan.so $GCC_LIB/libubsan.so" \
    ASAN_OPTIONS=detect_leaks=0:halt_on_error=0:log_path=$OUT/asan_oracle UBSAN_OPTIONS=print_stacktrace=1:log_path=$OUT/ubsan_oracle \
    python -m pytest tests/test_oracle_golden.py tests/test_oracle_vs_reference.py tests/test_fastmodel.py -q -p no:cacheprovider | tail -1
echo "reports that name libpmcts / liboracle (address / undefined / thread):"
grep -lE "libpmcts_|liboracle_|pmcts_oracle" "$OUT"/asan* "$OUT"/ubsan* "$OUT"/tsan* 2>/dev/null | while read -r f; do
    grep -E "^SUMMARY|runtime error" "$f" | grep -E "libpmcts_|liboracle_|pmcts_oracle" | sort | uniq -c
done
echo "(end of reports; logs under $OUT)"
