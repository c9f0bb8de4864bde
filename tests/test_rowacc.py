"""The hashed row accumulator of the assembly kernel (csrc/avs_rowacc.cuh, the default) against the linear one (AVS_ASM_ROW=linear),
compiled for the HOST: same entries, same insertion order, same bits -- for random rows, rows at the MAX_ROW limit, overflowing rows and keys that all collide in the table."""
import subprocess
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent

SRC = r'''
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>
#include "avs_rowacc.cuh"

static int compare(const std::vector<int32_t> &cols, const std::vector<double> &vals) {
    RowAcc a; RowAccHash h;
    a.init(); h.init();
    for (size_t i = 0; i < cols.size(); ++i) { a.add(cols[i], vals[i]); h.add(cols[i], vals[i]); }
    if (a.n != h.n || a.overflow != h.overflow) return 1;
    if (memcmp(a.col, h.col, sizeof(int32_t) * a.n) || memcmp(a.val, h.val, sizeof(double) * a.n)) return 2;
    return 0;
}

int main() {
    std::mt19937_64 rng(12345);
    long cases = 0;
    for (int trial = 0; trial < 20000; ++trial) {
        int distinct = 1 + (int)(rng() % 100);             // up to 100 distinct columns: beyond MAX_ROW = 80 overflows both alike
        int adds = 1 + (int)(rng() % 200);
        int32_t base = (int32_t)(rng() % 2000000000u);
        int stride = (trial % 3 == 0) ? 128 : (trial % 3 == 1 ? 1 : (int)(1 + rng() % 5000));   // stride 128 * k: worst-case clustering
        std::vector<int32_t> cols; std::vector<double> vals;
        for (int i = 0; i < adds; ++i) {
            cols.push_back((int32_t)(((long long)base + (long long)(rng() % distinct) * stride) % 2147483000LL));
            vals.push_back((double)(int64_t)(rng() % 2000001) / 1000.0 - 1000.0);
        }
        int rc = compare(cols, vals);
        if (rc) { printf("MISMATCH trial %d rc %d\n", trial, rc); return 1; }
        ++cases;
    }
    // keys that hash to the same slot: multiples of 2^32 / 128 stepped through the multiplicative hash's period
    {
        std::vector<int32_t> cols; std::vector<double> vals;
        for (int i = 0; i < MAX_ROW; ++i) { cols.push_back(i * 128 * 4096); vals.push_back(i + 0.5); }
        for (int i = 0; i < MAX_ROW; ++i) { cols.push_back(i * 128 * 4096); vals.push_back(0.25); }
        if (compare(cols, vals)) { printf("MISMATCH colliding keys\n"); return 1; }
    }
    printf("ok %ld\n", cases);
    return 0;
}
'''


def test_hashed_row_accumulator_is_bit_identical_to_the_linear_one(tmp_path):
    src = tmp_path / "rowacc.cpp"
    src.write_text(SRC)
    exe = tmp_path / "rowacc"
    subprocess.run(["/usr/bin/g++", "-O2", "-std=c++17", "-I", str(ROOT / "adaptiveviscositysolver_b200" / "csrc"), str(src), "-o", str(exe)],
                   check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.startswith("ok 20000")
