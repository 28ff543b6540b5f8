"""The boundary is a plain C ABI: a C program that includes include/psb200.h, links nothing but libdl and
loads libpsb200.so must be able to use it (this is what a ccall / cgo / JNI binding relies on)."""
import os
import subprocess
import textwrap

import pytest

from conftest import ROOT

C_SRC = textwrap.dedent(r"""
    #include <dlfcn.h>
    #include <stdio.h>
    #include <string.h>
    #include "psb200.h"

    typedef int (*mcm_t)(int, int, int, const double*, int, double*, long, double*, int);
    typedef int (*edges_t)(int, int, int, int, int*);
    typedef long long (*terms_t)(int, int, int, int);
    typedef const char* (*str_t)(void);
    typedef int (*cnt_t)(void);
    typedef int (*m2a_t)(int, int, int, int, const double* const*, double, double*);

    int main(int argc, char** argv)
    {
        void* h = dlopen(argv[1], RTLD_NOW | RTLD_LOCAL);
        if (!h) { printf("dlopen failed: %s\n", dlerror()); return 2; }
        mcm_t mcm = (mcm_t)dlsym(h, "psb200_mcm");
        edges_t edges = (edges_t)dlsym(h, "psb200_band_edges");
        terms_t terms = (terms_t)dlsym(h, "psb200_terms");
        str_t version = (str_t)dlsym(h, "psb200_version");
        str_t last_error = (str_t)dlsym(h, "psb200_last_error");
        cnt_t ndev = (cnt_t)dlsym(h, "psb200_device_count");
        m2a_t map2alm = (m2a_t)dlsym(h, "psb200_map2alm");
        if (!mcm || !edges || !terms || !version || !last_error || !ndev || !map2alm) { printf("missing symbol\n"); return 3; }
        printf("version=%s devices=%d\n", version(), ndev());
        int e[5];
        if (edges(0, 767, 768, 4, e) != 0 || e[0] != 0 || e[4] != 768) { printf("band_edges wrong\n"); return 4; }
        if (terms(1, 767, 0, 768) != 151289984LL) { printf("terms wrong\n"); return 5; }
        double V[8] = {1, 1, 1, 1, 1, 1, 1, 1}, M[64];
        memset(M, 0, sizeof M);
        int rc = mcm(PSB200_MPP_MMM, 0, 7, V, 8, M, 8, NULL, 1);      /* fused kind without the 2nd output */
        if (rc != 1) { printf("expected bad-argument code, got %d\n", rc); return 6; }
        rc = mcm(PSB200_M00, 0, 7, V, 8, M, 8, NULL, 1);
        if (ndev() == 0) {
            if (rc != 5 || !strstr(last_error(), "no CPU fallback")) { printf("expected code 5, got %d (%s)\n", rc, last_error()); return 7; }
        } else if (rc != 0) { printf("compute failed: %d %s\n", rc, last_error()); return 8; }
        /* W-spectrum production: the pixel-weighted analysis (niter 0) of 2 x (0.5 everywhere) at nside 2 gives a_00 = sqrt(4 pi) */
        double map[48], alm[2 * 10];
        for (int i = 0; i < 48; ++i) map[i] = 0.5;
        const double* maps[1] = {map};
        int rs = map2alm(2, 3, 0, 1, maps, 2.0, alm);
        if (ndev() == 0) {
            if (rs != 5) { printf("map2alm: expected code 5, got %d\n", rs); return 11; }
        } else if (rs != 0 || !(alm[0] > 3.5449077018110 && alm[0] < 3.5449077018111) || alm[1] != 0.0) {
            printf("map2alm failed: %d a00 = %.17g (%s)\n", rs, alm[0], last_error()); return 12;
        }
        if (map2alm(3, 3, 3, 1, maps, 1.0, alm) != 1) { printf("map2alm: nside 3 must be a bad argument\n"); return 13; }
        if (argc > 2 && !strcmp(argv[2], "--require-gpu")) {
            /* V == 1 and l1 + l2 <= 7 = nV-1: the l3 sum is complete, Xi = 1/4pi, M[l1,l2] = (2 l2+1)/4pi */
            if (rc != 0) { printf("a device is required here (rc=%d)\n", rc); return 9; }
            const double q = 0.07957747154594767;
            for (int l1 = 0; l1 < 8; ++l1)
                for (int l2 = 0; l1 + l2 < 8; ++l2) {
                    const double want = (2 * l2 + 1) * q, got = M[l1 + 8 * l2];
                    if (!(got > want * (1 - 1e-13) && got < want * (1 + 1e-13))) { printf("M[%d,%d] = %.17g, want %.17g\n", l1, l2, got, want); return 10; }
                }
            printf("computed on the device\n");
        }
        printf("ok rc=%d\n", rc);
        return 0;
    }
""")


def _build(tmp_path):
    src = tmp_path / "abi.c"
    exe = tmp_path / "abi"
    src.write_text(C_SRC)
    subprocess.run(["/usr/bin/gcc", "-std=c11", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                    str(src), "-o", str(exe), "-ldl"], check=True)
    return exe


def test_c_program_uses_the_abi(tmp_path, ps):
    """CPU box: symbols resolve, host-side helpers answer, a compute call fails loudly with code 5."""
    out = subprocess.run([str(_build(tmp_path)), ps.LIB_PATH], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "ok rc=" in out.stdout


@pytest.mark.gpu
def test_c_program_computes_on_the_device(tmp_path, ps):
    """GPU box: the same plain-C client must get a correct matrix out of psb200_mcm."""
    out = subprocess.run([str(_build(tmp_path)), ps.LIB_PATH, "--require-gpu"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "computed on the device" in out.stdout
