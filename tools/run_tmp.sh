#!/bin/bash
python tools/pool_sweep.py cornel_box 100 21 22 23 24 25 2>&1 | grep -v "^$"
python tools/pool_sweep.py final_scene 32 22 23 24 2>&1 | grep -v "^$"
python tools/pool_sweep.py one_weekend 32 22 23 24 25 2>&1 | grep -v "^$"
