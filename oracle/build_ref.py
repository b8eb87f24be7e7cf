"""
ORACLE tooling (test infrastructure, NOT product code).

Compiles the reference's own device kernels -- the four OKL files, where they lie under
/root/reference -- for the CPU, the way OCCA's Serial backend runs them: every
@outer/@inner/@tile loop becomes a plain loop, @shared arrays become block-local arrays,
@kernel functions become extern "C" functions.  The translated text only ever exists in a
temporary directory; the sole output is oracle/_ref/libref_okl_{2,3}d.so (git-ignored,
travels to the GPU box).  No reference source is copied into the repository.

The result is the *reference itself* for SURVEY.md section 2.2 rows D1-D9, S1-S14, C1-C3,
M1-M4, and is what tests/test_oracle_pin.py pins oracle/kernels.c against.

The OCCA compile-time defines are supplied as the reference supplies them
(domain.tpp:337-340, subdomain.tpp:3880-3890): DType/EType=double, BLOCK_SIZE=128,
OCCA_TYPE=0, DIM=2|3.  POLY_DEGREE -- a per-ladder literal table in the reference -- is
pointed at a settable global so one library serves every ladder.
"""
import os
import re
import subprocess
import sys
import tempfile

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")

PREAMBLE = """
typedef double DType;
typedef double EType;
#define BLOCK_SIZE 128
#define OCCA_TYPE 0
static double g_poly_degree[64];
extern "C" void ref_set_poly_degree(const double *p, int n) { for (int i = 0; i < n; i++) g_poly_degree[i] = p[i]; }
#define POLY_DEGREE const DType *poly_degree = g_poly_degree
"""


def translate(text, prefix):
    text = re.sub(r";\s*@tile\(.*?\)\)", ")", text)
    text = re.sub(r";\s*@outer\)", ")", text)
    text = re.sub(r";\s*@inner\)", ")", text)
    text = text.replace("@shared ", "")
    text = re.sub(r"@kernel\s+void\s+(\w+)\s*\(", lambda m: 'extern "C" void %s_%s(' % (prefix, m.group(1)), text)
    assert "@" not in re.sub(r"/\*.*?\*/", "", text, flags=re.S), "untranslated OKL attribute left"
    return text


def build(verbose=True):
    if not os.path.isdir(REF):
        return False
    os.makedirs(OUT, exist_ok=True)
    with tempfile.TemporaryDirectory() as tmp:
        for dim in (2, 3):
            src = os.path.join(tmp, "ref_okl_%dd.cpp" % dim)
            with open(src, "w") as f:
                f.write("#define DIM %d\n" % dim)
                f.write(PREAMBLE)
                for okl, prefix in (("domain.okl", "domain"), ("subdomain.okl", "subdomain"),
                                    ("csr_matrix.okl", "csr"), ("math.okl", "math")):
                    with open(os.path.join(REF, okl)) as g:
                        f.write("\n// ---- %s ----\n" % okl)
                        f.write(translate(g.read(), prefix))
            out = os.path.join(OUT, "libref_okl_%dd.so" % dim)
            # -O2 as in the reference Makefile:37; no -ffast-math, no FMA contraction (x86-64 baseline)
            cmd = ["g++", "-O2", "-shared", "-fPIC", "-o", out, src]
            if verbose:
                print(" ".join(cmd))
            subprocess.check_call(cmd)
    return True


if __name__ == "__main__":
    ok = build()
    print("built" if ok else "reference not present: nothing built")
    sys.exit(0)
