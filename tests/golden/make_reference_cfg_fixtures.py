"""Regenerates tests/golden/reference_cfg/ from the reference tree (run in the build container, where /root/reference exists):
verbatim copies of the configuration files the reference ships as tests / examples, and the YIG known-answer case that lives
as two raw strings inside src/jams/test/interactions.h:255-751 (424 interactions per primitive cell -> 424 * 8^3 = 217 088).
These are INPUT fixtures (config text), not reference source code."""
import os
import re
import shutil

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_cfg")
os.makedirs(OUT, exist_ok=True)
shutil.copy(os.path.join(REF, "test", "test_exchange_symops.cfg"), OUT)
shutil.copy(os.path.join(REF, "examples", "bloch_domain_wall", "bloch_domain_wall.cfg"), OUT)
src = open(os.path.join(REF, "src", "jams", "test", "interactions.h")).read()
start = src.index("TEST_F(MockJamsTest, generate_interactions_yig)")
blocks = re.findall(r'R"\((.*?)\)"', src[start:], re.S)
cfg, exc = blocks[0], blocks[1]
open(os.path.join(OUT, "yig_8x8x8.cfg"), "w").write("# src/jams/test/interactions.h:256-323 (generate_interactions_yig), verbatim\n" + cfg.strip("\n") + "\n")
open(os.path.join(OUT, "yig_princep_exc.in"), "w").write(exc.lstrip("\n"))
print("wrote", sorted(os.listdir(OUT)))
