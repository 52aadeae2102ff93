"""Build the UNMODIFIED reference MuZero and EfficientZero ctrees (Cython + C++) into oracle/_ref/.

TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is imported by the product
package `lightzero_b200`; only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may use it.

Sources are compiled from where they lie under /root/reference (read-only); no
reference source is copied into this repository.  Only the Cython-generated
.cpp (written to a temp dir) and the final extension module (oracle/_ref/, which
is git-ignored but travels to the GPU box) are produced.

Reference files compiled:
  lzero/mcts/ctree/ctree_muzero/mz_tree.pyx (+ mz_tree.pxd)
  lzero/mcts/ctree/ctree_muzero/lib/cnode.cpp, cnode.h   (textually included by the .pxd)
  lzero/mcts/ctree/common_lib/cminimax.cpp, cminimax.h, utils.cpp
  lzero/mcts/ctree/ctree_efficientzero/ez_tree.pyx (+ .pxd, lib/cnode.cpp, cnode.h), same common_lib
Flags mirror the reference setup.py:58,89-91 (-std=c++11, distutils -O2 -DNDEBUG).
Quirk: cnode.h:6 includes "./../common_lib/cminimax.h", which only
resolves with -I <ctree_muzero> on the include path.
"""
import os
import shutil
import subprocess
import sys
import sysconfig
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REF = os.environ.get("LZ_REFERENCE", "/root/reference")
CTREE = os.path.join(REF, "lzero", "mcts", "ctree")


def ref_module_path(name: str = "mz_tree"):
    suffix = sysconfig.get_config_var("EXT_SUFFIX")
    if name == "mz_tree_rand0":
        return os.path.join(OUT, "rand0", "mz_tree" + suffix)
    return os.path.join(OUT, name + suffix)


def load_rand0():
    """The shimmed MuZero tree as a separate module object (it shares the init symbol PyInit_mz_tree with the plain build)."""
    import importlib.machinery
    import importlib.util
    path = ref_module_path("mz_tree_rand0")
    if not os.path.exists(path):
        return None
    loader = importlib.machinery.ExtensionFileLoader("mz_tree", path)
    spec = importlib.util.spec_from_file_location("mz_tree", path, loader=loader)
    mod = importlib.util.module_from_spec(spec)
    loader.exec_module(mod)
    return mod


# module name -> reference sub-directory.  ez_tree (EfficientZero) has no `deterministic` flag: cselect_child always
# draws `rand() % ties` (ctree_efficientzero/lib/cnode.cpp:691) after reseeding from the wall clock
# (common_lib/utils.cpp:12-26).  To make the UNMODIFIED sources reproducible the module is linked with
# oracle/rand_shim.c, whose hidden-visibility `rand()` returns 0: index 0 of the tie list is the first position
# attaining the exact maximum (cnode.cpp:676-688), i.e. exactly the `deterministic=True` rule of the MuZero tree.
# "mz_tree_rand0": the MuZero tree again, with the shim, in oracle/_ref/rand0/ (the module itself is still called mz_tree):
# its *_with_reuse entry points break ties with rand() only (cselect_root_child, cnode.cpp:637-641; cselect_child(..., false),
# :879), so pinning them needs the same trick.
MODULES = {"mz_tree": ("ctree_muzero", False), "ez_tree": ("ctree_efficientzero", True), "mz_tree_rand0": ("ctree_muzero", True)}


def build(force: bool = False, name: str = "mz_tree") -> str:
    """Returns the path of the built module, or '' if the reference tree is absent."""
    subdir, shim = MODULES[name]
    target = ref_module_path(name)
    if os.path.exists(target) and not force:
        return target
    src_dir = os.path.join(CTREE, subdir)
    if not os.path.isdir(src_dir):
        return ""
    import numpy
    os.makedirs(os.path.dirname(target), exist_ok=True)
    pyx = "mz_tree" if name == "mz_tree_rand0" else name
    with tempfile.TemporaryDirectory() as tmp:
        gen_cpp = os.path.join(tmp, pyx + ".cpp")
        # cython reads the .pyx/.pxd in place; only the generated C++ goes to tmp
        subprocess.check_call(
            [sys.executable, "-m", "cython", "--cplus", "-3", "-I", src_dir,
             os.path.join(src_dir, pyx + ".pyx"), "-o", gen_cpp])
        inc = sysconfig.get_paths()["include"]
        extra = []
        if shim:
            shim_o = os.path.join(tmp, "rand_shim.o")
            subprocess.check_call(["gcc", "-O2", "-fPIC", "-c", os.path.join(HERE, "rand_shim.c"), "-o", shim_o])
            extra = [shim_o]
        cmd = ["g++", "-std=c++11", "-O2", "-DNDEBUG", "-fPIC", "-shared", "-w",
               "-I", inc, "-I", numpy.get_include(),
               "-I", src_dir,                      # resolves "lib/cnode.cpp" from the pxd
               "-I", os.path.join(src_dir, "lib"),
               "-I", CTREE,                        # cnode.h:6 quirk: ./../common_lib
               gen_cpp] + extra + ["-o", target]
        # cnode.h includes "./../common_lib/cminimax.h" relative to <subdir>/lib -> <subdir>/common_lib
        # (does not exist). g++ then searches -I dirs: <subdir>/./../common_lib == ctree/common_lib.
        subprocess.check_call(cmd)
    return target


def build_all(force: bool = False):
    return [build(force, n) for n in MODULES]


if __name__ == "__main__":
    for p in build_all(force="--force" in sys.argv):
        print(p if p else "reference tree not present; nothing built")
