"""Builds the C restatement (oracle/c/d3p_oracle.c) with gcc into oracle/c/_build/libd3p_oracle.so.  Test
infrastructure only; `__graft_entry__.build()` runs this, tests/test_oracle_c.py loads the result."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_build", "libd3p_oracle.so")


def build():
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    src = os.path.join(HERE, "d3p_oracle.c")
    if not os.path.exists(OUT) or os.path.getmtime(OUT) < os.path.getmtime(src):
        subprocess.run(["gcc", "-O2", "-std=c11", "-shared", "-fPIC", "-Wall", "-o", OUT, src], check=True)
    return OUT


if __name__ == "__main__":
    print(build())
