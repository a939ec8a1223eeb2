#!/bin/bash
# GRU scan variants side by side (CUDA events, scripts/prof_gru.py): v3 = register-resident FFMA (default), mma = mma.sync 3xTF32,
# ffma = first-generation scans.
for m in v3 mma ffma; do echo "== REFIL_GRU_MODE=$m"; REFIL_GRU_MODE=$m python scripts/prof_gru.py; done
