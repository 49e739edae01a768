#!/usr/bin/env bash
# Batch harness: runs a table of training jobs one after the other on the ranks of a cluster (default: one rank per GPU of
# this box). One line of the table = one job: experiment, aggregation rule, n, f, per-worker batch size, number of steps.
# Every job tees its output to E=<experiment>-R=<rule>-N=<n>-F=<f>-B=<batch>.{stdout,stderr} in the working directory;
# evaluation and summaries every 1000 steps, a checkpoint every 10 minutes. Ctrl-C stops the current job cleanly (the
# runner checkpoints on SIGINT) and skips the rest.
#
#   CLUSTER='{"ps": ["hostA:7000"], "workers": ["hostA:7001", "hostB:7001"]}' ./experiments.sh      # explicit cluster
#   ./experiments.sh my_jobs.txt                                                                     # table from a file

set -u
cluster="${CLUSTER:-local}"
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
logs="$(pwd)"
child=""

stop_current() {
  [ -n "${child}" ] && kill -INT "${child}" 2>/dev/null && wait "${child}" 2>/dev/null
  exit 130
}
trap stop_current INT TERM

jobs_table() {
  if [ $# -ge 1 ]; then
    cat "$1"
  else
    cat <<'TABLE'
# experiment  rule     n  f  batch  steps
mnist         average  2  0  50     100000
TABLE
  fi
}

jobs_table "$@" | while read -r experiment rule n f batch steps; do
  case "${experiment}" in ""|\#*) continue ;; esac
  tag="E=${experiment}-R=${rule}-N=${n}-F=${f}-B=${batch}"
  options=(--experiment "${experiment}" --experiment-args "batch-size:${batch}" --aggregator "${rule}"
           --nb-workers "${n}" --nb-decl-byz-workers "${f}" --max-step "${steps}"
           --evaluation-delta 1000 --evaluation-period -1 --summary-delta 1000 --summary-period -1
           --checkpoint-delta -1 --checkpoint-period 600
           --stdout-to "${logs}/${tag}.stdout" --stderr-to "${logs}/${tag}.stderr" --use-gpu --reuse-gpu --no-wait)
  echo "[experiments] ${tag}: ${steps} step(s) on cluster ${cluster}"
  python3 "${here}/deploy.py" --cluster "${cluster}" --deploy --runner "${options[*]}" < /dev/null &
  child=$!
  wait "${child}" || echo "[experiments] ${tag} ended with status $?"
  child=""
done
