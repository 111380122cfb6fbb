#!/bin/bash
# Descriptor-shift probe (tools/desc_shift_probe.cu): aligned starts first, then the two base_offset forms in separate processes.
mkdir -p gpurun_out
{ echo "## aligned starts, base_offset 0"; timeout 8 ./tools/build/desc_shift_probe 0 0 8 16 96; echo "rc=$?";
  echo "## base_offset = (start >> 7) & 7"; timeout 8 ./tools/build/desc_shift_probe 1; echo "rc=$?";
  echo "## base_offset = 0"; timeout 8 ./tools/build/desc_shift_probe 0; echo "rc=$?"; } > gpurun_out/r4f_desc_shift.log 2>&1
cat gpurun_out/r4f_desc_shift.log
