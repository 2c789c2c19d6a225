"""Mechanical evidence that the kernels are what DESIGN.md says they are: the mnemonics that only sm_100a tensor-core /
cluster / async-copy code produces must be present in the compiled objects (no GPU needed: cuobjdump on the in-tree
build products, see /opt/skills/guides/B200_PROFILING.md for the mnemonic list)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "clstm_b200", "build")


def sass(obj):
    path = os.path.join(BUILD, obj)
    if not os.path.exists(path):
        import clstm_b200
        clstm_b200.build()
    return subprocess.check_output(["cuobjdump", "-sass", path], text=True)


pytestmark = pytest.mark.skipif(shutil.which("cuobjdump") is None, reason="cuobjdump not installed")


def test_dense_products_use_tcgen05_with_tmem():
    s = sass("gemm_tc.o")
    assert "sm_100a" in s
    assert s.count("UTCHMMA") >= 24          # tcgen05.mma kind::tf32: 3 MMAs (3xTF32) x 4 K-slices per k-block, 2+ kernels
    assert "LDTM" in s                        # tcgen05.ld: accumulators come back from tensor memory
    assert "UTCBAR" in s                      # tcgen05.commit -> mbarrier
    assert "HMMA" not in s.replace("UTCHMMA", "")   # no legacy mma.sync path


def test_recurrent_kernels_use_packed_fp32_and_async_staging():
    s = sass("lstm.o")
    assert s.count("FFMA2") >= 200            # fma.rn.f32x2 chains of the register-resident kernels
    assert "LDGSTS" in s                      # cp.async staging of the streamed operands
    assert "MUFU.EX2" in s and "MUFU.RCP" in s


def test_cluster_kernels_use_dsmem_async_stores_and_mbarriers():
    s = sass("lstm_cluster.o")
    assert "STAS" in s                        # st.async into a peer CTA's shared memory
    assert "SYNCS.ARRIVE.TRANS64" in s and "TRYWAIT" in s     # mbarrier expect-tx / try_wait
    assert "UCGABAR_ARV" in s                 # barrier.cluster (prologue / epilogue, A/B variants)
    assert s.count("FFMA2") >= 400


def test_fused_peer_update_uses_system_scope_flags():
    s = sass("misc.o")
    assert ".SYS" in s                        # st.release.sys / ld.acquire.sys on the NVLink peer flags


def test_batched_recurrence_is_tcgen05_with_tma_and_tmem():
    # the north-star formulation of the recurrence (lstm_tc.cu): tcgen05.mma on TMA-staged operands, TMEM accumulators,
    # cluster multicast of the shared h tile, release-increments for the per-step exchange
    s = sass("lstm_tc.o")
    assert "sm_100a" in s
    assert s.count("UTCHMMA") >= 30           # forward (3 widths): 2 MMAs x 4 k slices per chunk; backward: 3 x 4 per chunk
    assert s.count("UTMALDG") >= 20           # cp.async.bulk.tensor: weight slices, h tiles (2-D multicast and 3-D), backward weight chunks
    assert "UTMALDG.3D" in s and "UTMALDG.2D.MULTICAST" in s
    assert "LDTM" in s and "UTCBAR" in s      # tcgen05.ld / tcgen05.commit
    assert "REDG.E.ADD.STRONG.GPU" in s       # red.release.gpu on the step counters
    assert "HMMA" not in s.replace("UTCHMMA", "")
    assert "FENCE.VIEW.ASYNC" in s            # proxy fences between generic-proxy writes and TMA / tensor-core reads


def test_cluster_resident_recurrence_keeps_weights_in_tensor_memory():
    # lstm_tcx.cu: tcgen05.mma with the A operand (the weight slice) in TMEM, written there once by tcgen05.st; h and the partial
    # sums travel between the CTAs of the cluster by bulk copies shared::cta -> shared::cluster; no cluster-scope acquire (which
    # compiles to an L1 invalidate) inside the step loops
    s = sass("lstm_tcx.o")
    assert "sm_100a" in s
    assert s.count("UTCHMMA tmem") >= 8 and "STTM" in s      # A-from-TMEM MMAs (the first operand after D is tmem[..]) + tcgen05.st
    assert "UBLKCP.S.S" in s                                  # cp.async.bulk shared -> peer shared memory
    assert "LDTM" in s and "UTCBAR" in s
    assert "UTMALDG" in s                                     # shared-memory form of the weight slice (A/B runs) is TMA-loaded
    assert s.count("CCTL.IVALL") <= 12                        # only the cluster barriers at group boundaries invalidate L1
    assert "HMMA" not in s.replace("UTCHMMA", "")


def test_persistent_gemm_is_tma_fed_tcgen05():
    s = sass("gemm_x.o")
    assert "sm_100a" in s
    assert s.count("UTCHMMA") >= 12           # 3 MMAs (hi*hi + hi*lo + lo*hi) x 4 k slices per 64-wide k block
    assert s.count("UTMALDG") >= 4            # A hi / lo and B hi / lo tiles by TMA
    assert "LDTM" in s and "UTCBAR" in s
    assert "HMMA" not in s.replace("UTCHMMA", "")
