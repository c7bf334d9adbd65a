#!/bin/bash
bash tools/ab_libs.sh "final_scene:32 two_perlin_spheres:32:1200" librt1w variant_p2 librt1w variant_p2
