#!/bin/bash
bash tools/ab_libs.sh "stress:8 final_scene:32 random_scene:32:1200 cornel_box:100 one_weekend:32" variant_psin librt1w
