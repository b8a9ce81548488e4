#!/bin/bash
# The shared cell / traceback / device-graph headers of the product (csrc/poa_cell.h, poa_dgraph.h) and the test-side host graph under
# AddressSanitizer + UBSan, driven by the CPU emulation tests.  Run from the repo root.
set -e
B=tests/emul/_build; T=$(mktemp -d)
cp $B/libpoa_emul.so $B/libmisscore_emul.so $T/
F="-O1 -g -std=c++17 -fPIC -shared -fsanitize=address,undefined -fno-omit-frame-pointer -w"
g++ $F tests/emul/poa_emul.cpp tests/emul/dgraph_emul.cpp tests/emul/poa_graph.cpp -o $B/libpoa_emul.so
g++ $F tests/emul/misscore_emul.cpp -o $B/libmisscore_emul.so
LD_PRELOAD=$(gcc -print-file-name=libasan.so):$(gcc -print-file-name=libubsan.so) \
ASAN_OPTIONS=detect_leaks=0:halt_on_error=1 UBSAN_OPTIONS=print_stacktrace=1:halt_on_error=1 \
python -m pytest tests/test_host_logic.py tests/test_misscore_cpu.py tests/test_dgraph_cpu.py -x -q -k "emul or pruned or lookahead or kernel or selfcheck or dgraph" || RC=$?
cp $T/libpoa_emul.so $T/libmisscore_emul.so $B/; touch $B/*.so
exit ${RC:-0}
