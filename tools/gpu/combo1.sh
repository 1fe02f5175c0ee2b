bash tools/gpu/quick.sh f1 "roi_align or extractor or bucket or smoke or golden or full_size"
bash tools/gpu/paste_exp.sh
