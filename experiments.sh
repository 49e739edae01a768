#! /bin/bash
# Batch harness (role of the reference's experiments.sh): start the ranks of the cluster with deploy.py, run a list of
# experiments one after the other, each with its own stdout/stderr files named E=<exp>-R=<gar>-N=<n>-F=<f>-B=<batch>.

CLUSTER_DEF=${CLUSTER_DEF:-local}   # JSON cluster specification, or 'local' = one worker rank per visible GPU of this box
RUNNING_PID=0

function run {
	local NAME=E=${1}-R=${2}-N=${3}-F=${4}-B=${5}
	python3 deploy.py --cluster "${CLUSTER_DEF}" --deploy --runner "\
		--experiment ${1} \
		--aggregator ${2} \
		--nb-workers ${3} \
		--nb-decl-byz-workers ${4} \
		--experiment-args batch-size:${5} \
		--max-step ${6} \
		--stdout-to ${NAME}.stdout \
		--stderr-to ${NAME}.stderr \
		--evaluation-period -1 \
		--checkpoint-period 600 \
		--summary-period -1 \
		--evaluation-delta 1000 \
		--checkpoint-delta -1 \
		--summary-delta 1000 \
		--use-gpu --reuse-gpu \
		--no-wait"&
	RUNNING_PID=$!
	trap run_abort TERM INT
	wait ${RUNNING_PID}
}

function run_abort {
	kill -s 2 ${RUNNING_PID}
	wait ${RUNNING_PID}
	exit 0
}

# Begin experiments
run mnist average 2 0 50 100000
# End experiments
