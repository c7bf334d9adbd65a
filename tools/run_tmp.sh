#!/bin/bash
bash tools/ab_libs.sh "final_scene:32 random_scene:32:1200 one_weekend:32" librt1w variant_splitsph variant_st8 variant_st12
