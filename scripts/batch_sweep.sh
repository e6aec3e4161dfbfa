#!/bin/bash
# L2-resident sub-batching probe: the same step at smaller batches; per-utterance stage times.
set -u
mkdir -p gpurun_out
for b in 4 8 16 32 64; do
  timeout 300 python bench.py --batch $b --steps $((512 / b)) --warmup 3 --no-e2e --no-cpu-baseline --no-alt-precision --no-config5 > gpurun_out/r2_batch_$b.json 2> gpurun_out/r2_batch_$b.err || { echo "batch $b failed"; tail -3 gpurun_out/r2_batch_$b.err; }
done
python - <<'PY'
import json
rows = {}
for b in (4, 8, 16, 32, 64):
    try:
        d = json.load(open(f'gpurun_out/r2_batch_{b}.json'))
    except Exception as e:
        print(b, 'missing', e); continue
    st = {s['kernel']: s['ms_per_step'] / b for s in d['stages']}
    rows[b] = dict(audio_s_per_s=d['value'], ms_per_utt=d['ms_per_step'] / b, gemm=st.get('gemm_bf16_tcgen05'), attn=st.get('attention_encoder'),
                   layernorm=st.get('layernorm'), sm_mhz=d['clocks']['sm_mhz'])
    print(b, {k: (round(v, 4) if isinstance(v, float) else v) for k, v in rows[b].items()})
json.dump(rows, open('gpurun_out/r2_batch_sweep.json', 'w'), indent=1)
PY
