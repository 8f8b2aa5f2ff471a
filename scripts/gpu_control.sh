#!/usr/bin/env bash
# MegaScan slow-rank injector: lock / reset / monitor SM clocks of one GPU (needs root).
#   gpu_control.sh limit <gpu> <MHz> | reset [gpu] | monitor [gpu]
# Without root use the software injector instead: pretrain_*.py --slow-rank-inject <rank>:<cycles>
set -euo pipefail
cmd="${1:-}"
case "$cmd" in
  limit)
    gpu="${2:?gpu index}"; mhz="${3:?MHz}"
    nvidia-smi -i "$gpu" -lgc "0,${mhz}"
    ;;
  reset)
    if [ -n "${2:-}" ]; then nvidia-smi -i "$2" -rgc; else nvidia-smi -rgc; fi
    ;;
  monitor)
    if [ -n "${2:-}" ]; then nvidia-smi dmon -i "$2" -s pc; else nvidia-smi dmon -s pc; fi
    ;;
  *)
    echo "usage: $0 {limit <gpu> <MHz>|reset [gpu]|monitor [gpu]}" >&2; exit 2;;
esac
