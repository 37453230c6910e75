#!/bin/bash
# cycle profile of the 16-lane scenes (development probe)
export TSIM_B200_LIB=$PWD/tactilesimulation_b200/_variants/prof16.so
PCASE=dclaw8x6_episodic_s0 PLANES=16 PB=2048 PT=200 timeout 600 python tools/cycle_profile.py 2>&1 | tee gpurun_out/q_dclaw.txt
PCASE=insertion20x20_episodic_s0 PLANES=16 PB=1024 PT=45 timeout 600 python tools/cycle_profile.py 2>&1 | tee gpurun_out/q_insertion.txt
