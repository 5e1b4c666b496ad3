#!/bin/bash
cd "$GRAFT_REPO_ROOT/scripts/probe"; mkdir -p ../../gpurun_out
for v in $(seq 0 7); do timeout 60 ./umma_probe_bf16 $v 2>&1 | tail -2; done | tee ../../gpurun_out/umma_probe_bf16.log
