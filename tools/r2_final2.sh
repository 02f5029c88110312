#!/bin/bash
set -u
tag=${1:-r2za}; out=gpurun_out; mkdir -p $out
rm -f $out/parity_report.jsonl
timeout 1800 python -m pytest tests -q -m gpu > $out/${tag}_pytest_all.log 2>&1; echo "pytest all exit $?"; tail -3 $out/${tag}_pytest_all.log
cp $out/parity_report.jsonl $out/${tag}_parity_report.jsonl 2>/dev/null
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1; echo "smoke exit $?"; tail -1 $out/${tag}_smoke.log
bash tools/r2_final.sh $tag
timeout 600 python tools/spconv_microbench.py --levels > $out/${tag}_micro_levels.txt 2>&1; grep -A3 "^L0" $out/${tag}_micro_levels.txt | cut -c1-110
