#!/bin/bash
# config 4 on N GPUs through the real corpus driver (strong scaling): bash scripts/r2_corpus.sh N [utts]
set -u
N=${1:-1}; U=${2:-16384}
mkdir -p gpurun_out
if [ "$N" = "1" ]; then
  timeout 900 python bench.py --workload corpus --utts $U > gpurun_out/r2_corpus_n${N}.json 2> gpurun_out/r2_corpus_n${N}.err
else
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --workload corpus --utts $U > gpurun_out/r2_corpus_n${N}.json 2> gpurun_out/r2_corpus_n${N}.err
fi
echo "corpus N=$N exit $?"; tail -c 600 gpurun_out/r2_corpus_n${N}.err; python - <<PY
import json
d=json.load(open('gpurun_out/r2_corpus_n${N}.json'))
print({k:d[k] for k in ('value','windows_audio_s_per_s','utt_per_s','ms_total','rank_ms','gather_ms','writer_ms_total','host_ms_per_batch_max_over_ranks','complete')})
PY
