#!/bin/bash
bash tools/gpu_multi.sh
bash tools/gpu_sweep.sh multi
