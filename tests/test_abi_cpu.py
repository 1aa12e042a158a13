"""CPU-only checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/mineral_b200.h declares, the ctypes table covers them all, and the struct layouts that
cross the ABI match what a C compiler sees.  No compute calls (no GPU here)."""

import ctypes
import os
import re
import subprocess
import tempfile

import pytest

from conftest import ROOT

HEADER = os.path.join(ROOT, "include", "mineral_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mnr_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from mineral_b200 import _native as nv

    names = declared_symbols()
    assert len(names) >= 29
    lib = ctypes.CDLL(nv.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
    assert sorted(nv.SIGNATURES) == names, "ctypes signature table and header disagree"
    assert nv.lib.mnr_abi_version() == 1
    assert nv.lib.mnr_error_string(-5).decode() == "workspace too small"
    assert nv.lib.mnr_sm_count() > 0


def test_struct_layouts_match_c():
    from mineral_b200 import _native as nv

    prog = r'''
#include <stdio.h>
#include "mineral_b200.h"
int main(void) {
  printf("%zu %zu %zu %zu ", sizeof(mnr_fold_seg_t), sizeof(mnr_fold_job_t), sizeof(mnr_heads_metrics_job_t), sizeof(mnr_fold_list_t));
  printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(mnr_wide_field_t), sizeof(mnr_row_field_t), sizeof(mnr_nstep_in_t),
         sizeof(mnr_nstep_out_t), sizeof(mnr_ppo_loss_cfg_t), sizeof(mnr_optim_state_t), sizeof(mnr_linear_group_t),
         sizeof(mnr_split_job_t), sizeof(mnr_ln_group_t), sizeof(mnr_critic_set_t), sizeof(mnr_peer_comm_t),
         sizeof(mnr_tc_reduce_job_t), sizeof(mnr_squash_cfg_t));
  return 0;
}'''
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "s.c")
        open(c, "w").write(prog)
        exe = os.path.join(d, "s")
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe], check=True)
        sizes = [int(x) for x in subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()]
    mine = [ctypes.sizeof(x) for x in (nv.FoldSeg, nv.FoldJob, nv.HeadsMetricsJob, nv.FoldList, nv.WideField, nv.RowField, nv.NStepIn, nv.NStepOut, nv.PPOLossCfg, nv.OptimState, nv.LinearGroup,
                                       nv.SplitJob, nv.LnGroup, nv.CriticSet, nv.PeerCommStruct, nv.TcReduceJob, nv.SquashCfg)]
    assert sizes == mine


def test_argument_validation_without_gpu():
    """Negative return codes come from host-side validation, before any launch."""
    from mineral_b200._native import lib

    assert lib.mnr_gae_f32(None, None, None, None, None, None, 4, 4, 0.99, 0.94, None) == -1  # MNR_ENULL
    one = ctypes.c_void_p(16)
    assert lib.mnr_gae_f32(one, one, one, one, one, None, 0, 4, 0.99, 0.94, None) == -2  # MNR_ESHAPE
    assert lib.mnr_critic_targets_f32(one, one, one, None, one, 4, 4, 7, 0.99, 0.95, 0.05, 0.94, 0.0, 0, 1.0, 1.0, None) == -4
    assert lib.mnr_post_rollout_workspace_bytes(32, 4096) > 0
    assert lib.mnr_rms_moments_workspace_bytes(131072, 37, 32768) > 0
    # deferred folds: registering empties the list; a reduce launch with neither jobs nor folds, or with a malformed fold, is refused
    from mineral_b200 import _native as nv
    fl = nv.FoldList()
    fl.n_folds, fl.has_metrics = 3, 1
    assert lib.mnr_fold_defer(ctypes.byref(fl)) == 0 and fl.n_folds == 0 and fl.has_metrics == 0
    assert lib.mnr_fold_defer(None) == 0
    assert lib.mnr_tc_reduce_partials_folds(None, 0, None, None) == -2
    assert lib.mnr_tc_reduce_partials_folds(None, 0, ctypes.byref(fl), None) == -2   # empty list
    fl.n_folds = 1   # parts == NULL
    assert lib.mnr_tc_reduce_partials_folds(None, 0, ctypes.byref(fl), None) == -1
    fl.folds[0].parts, fl.folds[0].n_parts, fl.folds[0].width, fl.folds[0].n_segs = 16, 4, 8, 1
    fl.folds[0].seg[0].begin, fl.folds[0].seg[0].end, fl.folds[0].seg[0].dst = 0, 9, 16   # segment past the row width
    assert lib.mnr_tc_reduce_partials_folds(None, 0, ctypes.byref(fl), None) == -2


def test_product_refuses_cpu_tensors():
    """No CPU fallback: handing the product a CPU tensor is an error, not a slow path."""
    import torch

    from mineral_b200 import ops
    from mineral_b200._native import MineralB200Error

    T, N = 4, 8
    with pytest.raises(MineralB200Error):
        ops.gae(torch.zeros(T, N, 1), torch.zeros(T, N, 1), torch.zeros(T, N, dtype=torch.uint8), torch.zeros(N, 1), 0.99, 0.95)


def test_product_does_not_import_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "mineral_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("# oracle", ""), f"{f} mentions the oracle"
