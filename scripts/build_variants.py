"""Developer A/B builds of the CUDA library: python scripts/build_variants.py name:DEF1,DEF2 ...
-> seam-match-rcnn_b200/libseam_b200.<name>.so, selected at run time with SEAM_B200_LIB=<path>."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import seam_match_rcnn_b200 as pkg
for spec in sys.argv[1:]:
    name, _, defs = spec.partition(":")
    print(sys.modules[pkg.__name__ + "._build"].build_variant(name, [d for d in defs.split(",") if d]))
