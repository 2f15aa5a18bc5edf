#!/bin/bash
NS=4 STEPS=10 bash tools/r02/gpu_r2_n8.sh
