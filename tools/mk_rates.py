"""MK NAND gates/s for 2/4/8 parties with key-shaped random material (bench.py's mk_nand leg on its own; development A/B)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import tfhe_jl_b200 as T
from tfhe_jl_b200 import _cabi
print(json.dumps({"mk_pw": os.environ.get("TFHE_B200_MK_PW", "1"), "mk_nand": bench.mk_nand_rates(T, _cabi, torch, 0)}))
