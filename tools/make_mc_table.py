"""Regenerate commonscenes_b200/model/diff_utils/_mc_table.py from the construction in oracle/mesh.py (dev-time tool)."""
import os, sys, textwrap
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import mesh

cnt, tab = mesh.triangle_table()
print("MAX_TRIS =", tab.shape[1])
print("TRI_COUNT =", "\n".join(textwrap.wrap(cnt.tobytes().hex(), 112)))
print("TRI_TABLE =", "\n".join(textwrap.wrap(tab.tobytes().hex(), 112)))
