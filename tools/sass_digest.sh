#!/bin/bash
# tools/sass_digest.sh > profiles/r2_sass_digest.txt : opcode histogram of the fused kernels in libscipnp.so (cuobjdump -sass),
# the evidence that the hot path is TMA (UTMALDG/UTMASTG), mbarrier (SYNCS) and packed-FP32 (FFMA2/FMUL2/FADD2) code.
cd "$(dirname "$0")/../sci-algorithms_b200"
for obj in ws_inst_r4 fused_inst_r4; do
  echo "== build/$obj.o (nvcc -gencode arch=compute_100a,code=sm_100a), all template instances"
  cuobjdump -sass build/$obj.o | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed -E 's/^\s+\/\*[0-9a-f]+\*\/\s+(@!?U?P[0-9T]+ )?//' | awk '{print $1}' | sed -E 's/^([A-Z0-9]+(\.(SQRT|RCP|RSQ|128|64|UP|DOWN|BFLY|IDX|[0-9]D|TRANS64|TRYWAIT|ARRIVE|EXCH|PHASECHK))?).*/\1/' | sort | uniq -c | sort -rn | head -40
done
echo "== gap_tv_ws_kernel<R=4, GAP accelerated, C=24>: fast block of four rows (the steady-state loop body of a consumer warp)"
cuobjdump -sass build/ws_inst_r4.o | awk '/Function : .*gap_tv_ws_kernelILi4ELi0ELi12/{f=1} f{print} /Function : .*gap_tv_ws_kernelILi4ELi0ELi4E/{if(f)exit}' | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed -E 's/^\s+\/\*[0-9a-f]+\*\/\s+(@!?U?P[0-9T]+ )?//' | awk '{print $1}' > /tmp/_k.txt
python3 - <<'PY'
ops = [l.strip() for l in open('/tmp/_k.txt')]
# the fast block: eight LDS.128 (f(t) and f(t-R) of four rows) with no mbarrier instruction in between, up to the last STS.64 of the fourth row
for i, o in enumerate(ops):
    if o.startswith('LDS.128'):
        idx = [k for k in range(i, min(i + 1200, len(ops))) if ops[k].startswith('LDS.128')][:8]
        if len(idx) == 8 and not any(x.startswith('SYNCS') or x.startswith('BAR') for x in ops[i:idx[7]]):
            stop = next((k for k in range(idx[7], min(idx[7] + 400, len(ops))) if ops[k].startswith('SYNCS') or ops[k].startswith('BAR')), idx[7] + 400)
            end = max(k for k in range(idx[7], stop) if ops[k].startswith('STS'))
            blk = ops[i:end + 1]
            from collections import Counter
            c = Counter(x.split('.')[0] for x in blk)
            print('%d instructions for 4 rows (%.1f per row and warp):' % (len(blk), len(blk) / 4.))
            for k, v in c.most_common():
                print('%7d %s' % (v, k))
            break
PY
