#!/usr/bin/env bash
# CONF_FILE=<yaml stem> [COMPILE=0/1] [TRAIN_ITERS=n] ./train.sh
# COMPILE=1 keeps the reference's observable behaviour (10-step dry run, no checkpoints / TensorBoard);
# there is no ahead-of-time graph compile on B200: kernels are prebuilt by `python __graft_entry__.py`.
cd "$(dirname "$0")"
source ./train_setup.sh

: ${CONF_FILE:=hf_llama3_8B_config}
CONF_FILE_PATH="./conf/${CONF_FILE}.yaml"
if [ ! -f "$CONF_FILE_PATH" ]; then
    echo "Error: YAML file '$CONF_FILE_PATH' not found!" >&2
    exit 1
fi
PROCESSES_PER_NODE=$(awk '/^[[:space:]]*devices:/ {print $2; exit}' "$CONF_FILE_PATH")
export PROCESSES_PER_NODE=${PROCESSES_PER_NODE:-1}

DISTRIBUTED_ARGS="--nproc_per_node $PROCESSES_PER_NODE --nnodes $NTASKS --node_rank $NODEID --master_addr $MASTER_ADDR --master_port $MASTER_PORT"
echo "$DISTRIBUTED_ARGS"

python -m torch.distributed.run $DISTRIBUTED_ARGS training_orchestrator.py \
    --config-path=conf \
    --config-name=$CONF_FILE \
    trainer.devices=$PROCESSES_PER_NODE \
    trainer.num_nodes=$NTASKS \
    exp_manager.explicit_log_dir=$EXPLICIT_LOGDIR "$@" 2>&1 | tee -a "$LOG_PATH/log"
