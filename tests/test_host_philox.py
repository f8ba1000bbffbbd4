"""The counter-based generator of the kernels (sslrec_b200/csrc/common.cuh is __host__ __device__) compiled for the HOST and
compared bit for bit with the numpy restatement (oracle/philox.py): the product header and the oracle agree on every draw the
GPU tests then check in-kernel."""
import os
import shutil
import subprocess
import tempfile

import numpy as np
import pytest

from oracle import philox as P

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SRC = r'''
#include <cstdio>
#include <cstdint>
#include "common.cuh"
namespace ssl { void set_error(const char *, ...) {} void count_launch(int) {} }
int main() {
    const uint64_t seed = 0x1234567890ABCDEFull;
    for (uint32_t r = 0; r < 40; ++r)
        for (uint32_t c = 0; c < 25; ++c) {
            const uint4 x = ssl::philox4x32_10(make_uint4(r, c * 7919u, 3u, 0x45444745u), ssl::seed_key(seed));
            printf("P %u %u %u %u\n", x.x, x.y, x.z, x.w);
            printf("E %d %d\n", (int)ssl::edge_keep_rng(seed, 3u, r, c * 7919u, 0.37f), (int)ssl::edge_keep_rng(seed, 0u, c, r, 0.9f));
        }
    for (uint32_t r = 0; r < 500; ++r) printf("N %d\n", (int)ssl::node_keep_rng(seed, r, 0.5f));
    for (uint32_t r = 0; r < 20; ++r)
        for (uint32_t q = 0; q < 12; ++q) {
            const float4 u = ssl::noise_u4_rng(seed, 2u, r + 1000u, q);
            printf("U %.9g %.9g %.9g %.9g\n", u.x, u.y, u.z, u.w);
        }
    return 0;
}
'''


@pytest.mark.skipif(shutil.which('nvcc') is None, reason='needs nvcc (host-only compile; no GPU involved)')
def test_common_cuh_generator_matches_numpy_oracle_on_the_host():
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, 't.cu'), 'w').write(SRC)
        exe = os.path.join(d, 't')
        subprocess.run(['nvcc', '-std=c++17', '-O1', '-I', os.path.join(ROOT, 'sslrec_b200', 'csrc'), os.path.join(d, 't.cu'), '-o', exe],
                       check=True, capture_output=True)
        out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout.splitlines()
    seed = 0x1234567890ABCDEF
    rr, cc = np.meshgrid(np.arange(40), np.arange(25) * 7919, indexing='ij')
    words = P.philox4x32_10(rr.ravel(), cc.ravel(), 3, P.TAG_EDGE, seed)
    got_p = np.array([[int(v) for v in ln.split()[1:]] for ln in out if ln.startswith('P')], dtype=np.uint64)
    assert np.array_equal(got_p, np.stack(words, 1).astype(np.uint64))
    got_e = np.array([[int(v) for v in ln.split()[1:]] for ln in out if ln.startswith('E')], dtype=bool)
    assert np.array_equal(got_e[:, 0], P.edge_keep(seed, 3, rr.ravel(), cc.ravel(), 0.37))
    assert np.array_equal(got_e[:, 1], P.edge_keep(seed, 0, cc.ravel() // 7919, rr.ravel(), 0.9))
    got_n = np.array([int(ln.split()[1]) for ln in out if ln.startswith('N')], dtype=bool)
    assert np.array_equal(got_n, P.node_keep(seed, np.arange(500), 0.5)) and 0.4 < got_n.mean() < 0.6
    got_u = np.array([[float(v) for v in ln.split()[1:]] for ln in out if ln.startswith('U')], dtype=np.float32).reshape(20, 48)
    assert np.array_equal(got_u.view(np.uint32), P.noise_uniform(seed, 2, 20, 48, row_offset=1000).view(np.uint32))
