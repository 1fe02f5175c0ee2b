bash tools/gpu/quick.sh c3
bash tools/gpu/prof_ra.sh ra_v3b
