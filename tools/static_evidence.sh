#!/bin/bash
# Static (no GPU) evidence of the built library: registers / stack / spills per kernel from the ptxas logs the
# Makefile keeps, and the census of Blackwell-native SASS mnemonics (B200_PROFILING.md "What proves a Blackwell-native
# kernel").  Usage: tools/static_evidence.sh > profiles/rNN_static_evidence.txt   (after build())
cd "$(dirname "$0")/.."
echo "## ptxas: registers / stack / spill bytes per kernel (top_down_renderer_b200/csrc/build/*.ptxas.log)"
for f in top_down_renderer_b200/csrc/build/*.ptxas.log; do python3 - "$f" <<'PY'
import re, sys, os
t = open(sys.argv[1]).read()
pat = (r"Compiling entry function '([^']+)' for 'sm_100a'.*?\n.*?(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads"
       r"\n.*?Used (\d+) registers(?:, used (\d+) barriers)?(?:, (\d+) bytes smem)?")
for name, stack, st, ld, regs, bars, smem in re.findall(pat, t, re.S):
    print(f"{os.path.basename(sys.argv[1])[:-10]:15s} {name[:72]:72s} regs {regs:>3s}  stack {stack:>4s}  spill st/ld {st}/{ld}  static smem {smem or 0}")
PY
done
echo
echo "## SASS mnemonics in libtdr_b200.so (cuobjdump -sass): tcgen05.mma = UTC*MMA, tcgen05.ld/st = LDTM/STTM, bulk async copies = UBLKCP, mbarrier = SYNCS"
cuobjdump -sass top_down_renderer_b200/libtdr_b200.so | grep -oE "\b(UTC[A-Z]*MMA[A-Z.0-9_]*|UTCBAR[A-Z.0-9_]*|UTCATOMSWS[A-Z.0-9_]*|UTCCP[A-Z.0-9_]*|LDTM[A-Z.0-9_]*|STTM[A-Z.0-9_]*|UTMALDG[A-Z.0-9_]*|UTMASTG[A-Z.0-9_]*|UBLKCP[A-Z.0-9_]*|SYNCS[A-Z.0-9_]*|HMMA[A-Z.0-9_]*|HGMMA[A-Z.0-9_]*)" | sort | uniq -c | sort -rn
echo
echo "## kernels that contain UTCHMMA"
cuobjdump -sass top_down_renderer_b200/libtdr_b200.so | awk '/Function :/{fn=$3} /UTCHMMA/{c[fn]++} END{for (f in c) print c[f], f}' | sort -rn
