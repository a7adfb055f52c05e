#!/usr/bin/env python
"""Builds oracle/_ref/inject_ref*.so -- TEST INFRASTRUCTURE ONLY.

The reference's wall-injection routine has no C symbol: its body is Cython
(/root/reference/infMagSim_cython.py:230-285) and the module it lives in cannot be built here
(it imports cython_gsl at module scope, :17-18, and links GSL).  inject_particles itself uses no
GSL, so this script cuts exactly that function out of the reference file WHERE IT LIES:

  * the four import lines the function needs (`cimport cython`, `import numpy as np`,
    `cimport numpy as np`, `import math`; ref :11-14),
  * the `cdef extern` declarations of draw_velocities_c and sgenrand and the seedgenrand_c
    wrapper (ref :224-228),
  * `def inject_particles(...)` up to, not including, the next `cdef extern` (ref :230-286),

writes them UNCHANGED into a temporary .pyx outside the repository, compiles it with Cython in
`language_level=2` (the reference is Python 2: its two `print` statements are accepted as they
are -- no line of the function is edited) and links the result against the reference's own
unmodified infMagSim_c.c compiled with the reference's flags (Makefile:16-18 there).  Only the
binary lands in oracle/_ref/ (git-ignored); no reference source is copied into the repository.

Two extensions are built from the same unchanged lines:
  inject_ref        linked against the reference's infMagSim_c.c: inject_particles draws its
                    velocities with the reference's own draw_velocities_c / MT19937;
  inject_ref_given  linked against oracle/inject_shim.c instead, whose draw_velocities_c hands
                    out velocities the test supplied (inject_shim_set_velocities): the reference's
                    injection body on caller-given random inputs, for lockstep tests against the
                    CUDA path, whose sampler is a different generator.
oracle/cpu.py imports them as the "reference" arm of the injection parity tests and
tests/golden/make_golden.py generates the injection fixture from the first.

    python oracle/build_inject_ref.py          # needs /root/reference (build container only)
"""
import glob
import os
import re
import shutil
import subprocess
import sys
import sysconfig
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE = os.environ.get("REFERENCE", "/root/reference")
OUT_DIR = os.path.join(HERE, "_ref")
MODULE = "inject_ref"
REF_FLAGS = ["-pthread", "-fPIC", "-fwrapv", "-O2", "-fno-strict-aliasing", "-w"]


def extract(lines):
    """The lines of the reference file that make up inject_particles and what it needs."""
    header = [l for l in lines[:40] if re.match(r"^(cimport cython|import numpy as np|cimport numpy as np|import math)\s*$", l)]
    if len(header) != 4:
        raise RuntimeError(f"expected the 4 import lines near the top of the reference file, found {header}")
    start = next(i for i, l in enumerate(lines) if l.startswith("cdef extern void draw_velocities_c"))
    begin_def = next(i for i, l in enumerate(lines) if l.startswith("def inject_particles("))
    if not (start < begin_def):
        raise RuntimeError("reference layout changed: draw_velocities_c extern is not in front of inject_particles")
    end = next(i for i in range(begin_def + 1, len(lines)) if lines[i].startswith("cdef extern"))
    return header + ["\n"] + lines[start:end]


def built_path():
    hits = sorted(glob.glob(os.path.join(OUT_DIR, MODULE + "*.so")))
    return hits[0] if hits else None


def main() -> int:
    src = os.path.join(REFERENCE, "infMagSim_cython.py")
    csrc = os.path.join(REFERENCE, "infMagSim_c.c")
    if not (os.path.exists(src) and os.path.exists(csrc)):
        print(f"reference sources absent under {REFERENCE}; keeping prebuilt {built_path()}")
        return 0
    import numpy as np
    from Cython.Build import cythonize  # noqa: F401  (fail early when Cython is missing)
    with open(src) as fh:
        body = extract(fh.readlines())
    os.makedirs(OUT_DIR, exist_ok=True)
    tmp = tempfile.mkdtemp(prefix="inject_ref_")
    try:
        pyx = os.path.join(tmp, MODULE + ".pyx")
        with open(pyx, "w") as fh:
            fh.writelines(body)
        subprocess.check_call([sys.executable, "-m", "cython", "-2", "--directive", "language_level=2", pyx,
                               "-o", os.path.join(tmp, MODULE + ".c")])
        ext = sysconfig.get_config_var("EXT_SUFFIX")
        inc = [sysconfig.get_paths()["include"], np.get_include()]
        base = ["gcc", "-shared", *REF_FLAGS, "-DNPY_NO_DEPRECATED_API=NPY_1_7_API_VERSION", *[f"-I{d}" for d in inc]]
        out = os.path.join(OUT_DIR, MODULE + ext)
        subprocess.check_call(base + [os.path.join(tmp, MODULE + ".c"), csrc, "-o", out, "-lm"])
        print(f"built {out} from {src}:inject_particles + {csrc}")
        # second flavour: same lines, module name inject_ref_given, velocities supplied by the caller
        pyx2 = os.path.join(tmp, MODULE + "_given.pyx")
        shutil.copy(pyx, pyx2)
        subprocess.check_call([sys.executable, "-m", "cython", "-2", "--directive", "language_level=2", pyx2,
                               "-o", os.path.join(tmp, MODULE + "_given.c")])
        out2 = os.path.join(OUT_DIR, MODULE + "_given" + ext)
        subprocess.check_call(base + [os.path.join(tmp, MODULE + "_given.c"), os.path.join(HERE, "inject_shim.c"),
                                      "-o", out2, "-lm"])
        print(f"built {out2} from {src}:inject_particles + oracle/inject_shim.c")
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
