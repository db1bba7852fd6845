#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python tools/probe_pk2.py > gpurun_out/probe_pk2.log 2>&1; cat gpurun_out/probe_pk2.log | cut -c1-200
