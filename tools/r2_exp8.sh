#!/bin/bash
timeout 300 python tools/time_gn.py > gpurun_out/x8_time_gn.txt 2>&1; tail -n 12 gpurun_out/x8_time_gn.txt
