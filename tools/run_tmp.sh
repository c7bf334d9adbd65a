#!/bin/bash
bash tools/ab_libs.sh "cornel_box:100 cornel_smoke:32 final_scene:32" librt1w variant_opsu librt1w variant_opsu
