#!/usr/bin/env bash
# Cluster environment discovery: SLURM, OpenMPI (k8s/EKS style launchers) or a single node.
# Exports NTASKS, NODEID, MASTER_ADDR, MASTER_PORT, LOG_PATH, EXPLICIT_LOGDIR for train.sh.
ulimit -n 65535 2>/dev/null || true

if [ -n "$SLURM_NTASKS" ]; then
    # one task per node under srun
    export NTASKS=$SLURM_NTASKS
    export NODEID=$SLURM_NODEID
    export MASTER_ADDR=$(scontrol show hostnames "$SLURM_JOB_NODELIST" | head -n 1)
    export LOG_PATH=${LOG_PATH:-logs/$SLURM_JOB_ID/${SLURM_RESTART_COUNT:-0}/$NODEID}
    : ${NEMO_EXPM_VERSION:=$SLURM_JOB_ID}
elif [ -n "$OMPI_COMM_WORLD_SIZE" ]; then
    export NTASKS=$OMPI_COMM_WORLD_SIZE
    export NODEID=$OMPI_COMM_WORLD_RANK
    : ${MASTER_ADDR:=$(echo "${OMPI_MCA_orte_hnp_uri:-127.0.0.1}" | sed -e 's/.*tcp:\/\///' -e 's/[,:].*//')}
    export MASTER_ADDR
    export LOG_PATH=${LOG_PATH:-logs/ompi/$NODEID}
else
    # single node by default; a manual multi-node launch presets NTASKS / NODEID / MASTER_ADDR on every node
    export NTASKS=${NTASKS:-1}
    export NODEID=${NODEID:-0}
    export MASTER_ADDR=${MASTER_ADDR:-127.0.0.1}
    export LOG_PATH=${LOG_PATH:-logs/local}
fi
mkdir -p "$LOG_PATH"
export MASTER_PORT=${MASTER_PORT:-41000}
export NEMO_EXPM_VERSION=${NEMO_EXPM_VERSION:-version_0}
export EXPLICIT_LOGDIR=${EXPLICIT_LOGDIR:-null}
export MALLOC_ARENA_MAX=128
# NCCL over NVLink5/NVSwitch inside the node; keep async error handling on so a dead rank aborts the job
export TORCH_NCCL_ASYNC_ERROR_HANDLING=${TORCH_NCCL_ASYNC_ERROR_HANDLING:-1}
export NCCL_NVLS_ENABLE=${NCCL_NVLS_ENABLE:-1}
export PYTORCH_CUDA_ALLOC_CONF=${PYTORCH_CUDA_ALLOC_CONF:-expandable_segments:True}
