"""CPU checks of the C++ shadow headers (minife_b200/host) that need no device -- not gpu.  Compiled against the
reference's own type headers, so they only run where the reference tree is present (this container)."""
import os
import subprocess
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/ref"

PROGRAM = r"""
#include <iostream>
#include <string>
#include <Vector.hpp>
#include <Vector_functions.hpp>
int main(int argc, char **argv)
{
  miniFE::Vector<double, int, int> v(7, 5);
  double vals[5] = { 0.0, -0.0, 1.0 / 3.0, 1e-300, -123456789.125 };
  for (int i = 0; i < 5; ++i) v.coefs[i] = vals[i];
  miniFE::write_vector(argv[1], v);
  miniFE::Vector<double, int, int> e(0, 0);
  miniFE::write_vector(std::string(argv[1]) + "_empty", e);
  int idx[4] = { 6, 7, 11, 12 }; double c[4] = { 1, 2, 3, 4 };
  miniFE::sum_into_vector(4, idx, c, v);
  miniFE::write_vector(std::string(argv[1]) + "_sum", v);
  return 0;
}
"""


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "src", "Vector.hpp")), reason="reference tree not present")
def test_write_vector_and_sum_into_vector_match_the_reference_bytes(product_lib):
    """write_vector (the MINIFE_DEBUG dump format, ref/src/Vector_functions.hpp:49-82) and sum_into_vector (84-103) of the
    shadow header against the reference's own, on a Vector that is not bound to a device context."""
    inc_ref = ["-I%s/src" % REF, "-I%s/utils" % REF, "-I%s/fem" % REF]
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "t.cpp")
        open(src, "w").write(PROGRAM)
        subprocess.run(["g++", "-O1", "-w", "-I%s/minife_b200/host" % ROOT, "-I%s/include" % ROOT] + inc_ref + ["-o", os.path.join(d, "ours.x"), src,
                        "-L%s/minife_b200" % ROOT, "-lminife_b200", "-Wl,-rpath,%s/minife_b200" % ROOT], check=True)
        subprocess.run(["g++", "-O1", "-w"] + inc_ref + ["-o", os.path.join(d, "ref.x"), src], check=True)
        subprocess.run([os.path.join(d, "ours.x"), os.path.join(d, "o")], check=True)
        subprocess.run([os.path.join(d, "ref.x"), os.path.join(d, "r")], check=True)
        for suffix in ("", "_empty", "_sum"):
            a = open(os.path.join(d, "o%s.1.0" % suffix)).read()
            b = open(os.path.join(d, "r%s.1.0" % suffix)).read()
            assert a == b and (suffix == "_empty" or len(a) > 10), suffix
