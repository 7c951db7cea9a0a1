#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python tools/diag_m_parity.py > gpurun_out/diag_m_parity.log 2>&1; echo rc=$?; grep -v Warning gpurun_out/diag_m_parity.log | tail -8
