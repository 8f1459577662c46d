"""World-size-2 gloo tests of the image-sharded runner (the N>1 path has no data-path collective: only the final gather)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from image_editing_framework_b200 import runner

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _edit_worker(i):
    """A miniature edit through the real host logic (oracle-backed ops on CPU), seeded by the image index."""
    from contextlib import redirect_stdout
    import io
    from oracle import cpu_ops
    from image_editing_framework_b200 import masactrl, editing
    from image_editing_framework_b200.standin import make_pipeline, tiny_config
    with cpu_ops.patched(), redirect_stdout(io.StringIO()):
        pipe = make_pipeline(tiny_config(), seed=0)
        ed = masactrl.MutualSelfAttentionControl(1, 10, total_steps=2)
        masactrl.regiter_attention_editor_diffusers(pipe, ed)
        lat = torch.randn(1, 4, 8, 8, generator=torch.Generator().manual_seed(100 + i))
        out = editing.masactrl_edit(pipe, ["a cat", "a dog"], torch.cat([lat, lat]), num_inference_steps=2)
    return out


def _rank_main(rank, ws, port, n_items, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    try:
        torch.set_num_threads(2)
        assert runner.world() == (rank, ws)
        assert runner.shard_indices(n_items, rank, ws) == [i for i in range(n_items) if i % ws == rank]
        squares = runner.run_sharded(lambda i: i * i, n_items)
        edits = runner.run_sharded(_edit_worker, 3)
        q.put((rank, squares, {k: v.numpy().copy() for k, v in edits.items()}))  # by value: no shared-memory handles
    finally:
        dist.destroy_process_group()


def test_run_sharded_world2_matches_serial():
    n_items, ws = 7, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_rank_main, args=(r, ws, port, n_items, q)) for r in range(ws)]
    for p in procs:
        p.start()
    got = [q.get(timeout=300) for _ in range(ws)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    serial = {i: _edit_worker(i) for i in range(3)}
    for rank, squares, edits in got:
        assert squares == {i: i * i for i in range(n_items)}
        assert sorted(edits) == [0, 1, 2]
        for i in range(3):
            assert torch.equal(torch.from_numpy(edits[i]), serial[i]), f"rank {rank} item {i}: sharded result differs from the 1-process run"


def test_run_sharded_without_process_group_is_serial():
    assert runner.world() == (0, 1)
    assert runner.run_sharded(lambda i: -i, 4) == {0: 0, 1: -1, 2: -2, 3: -3}
    assert runner.shard_indices(10, 3, 4) == [3, 7]


def test_sweep_shards_reproduce_the_serial_sweep_and_retarget_equals_a_fresh_controller(monkeypatch):
    """tools/sweep.py (BASELINE configs[4]) on the CPU with the oracle-backed ops: the union of the 2-rank shards (image i -> rank
    i mod 2, kept controllers re-pointed with retarget()) carries exactly the per-image CRCs of the 1-rank sweep, replace and refine
    pairs both occur, and a retargeted controller's tables equal those of a freshly built one."""
    import importlib.util
    import torch
    from oracle import cpu_ops
    import image_editing_framework_b200 as pkg
    cpu_ops.install(monkeypatch)
    spec = importlib.util.spec_from_file_location("ief_sweep", os.path.join(ROOT, "tools", "sweep.py"))
    sweep = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(sweep)
    methods, n = ["p2p", "masactrl", "pnp"], 5
    make = lambda: sweep.Worker(torch.device("cpu"), 3, False, "tiny", torch.float32)
    serial = sweep.sweep(make(), n, methods, rank=0, world=1)
    sharded = {}
    for rank in range(2):
        sharded.update(sweep.sweep(make(), n, methods, rank=rank, world=2))
    assert sorted(sharded) == sorted(serial) == list(range(n))
    for i in range(n):
        assert sharded[i]["crc"] == serial[i]["crc"], i
    kinds = {len(r["source"].split(" ")) == len(r["target"].split(" ")) for r in serial.values()}
    assert kinds == {True, False}
    assert len({r["crc"]["p2p"] for r in serial.values()}) == n        # the images differ, so must the results
    tok = pkg.standin.WordPieceTokenizer()
    a, b = ["a photo of a cat on a bench", "a photo of a dog on a bench"], ["a sketch of a fox by the lake", "a sketch of a squirrel by the lake"]
    kept, fresh = pkg.p2p.AttentionReplace(a, tok, 5, 0.8, 0.6, device="cpu"), pkg.p2p.AttentionReplace(b, tok, 5, 0.8, 0.6, device="cpu")
    kept._edit = kept.cross_edit()
    kept.retarget(b, tok)
    assert torch.equal(kept.mapper, fresh.mapper) and torch.equal(kept._alpha_table, fresh._alpha_table)
    e = fresh.cross_edit()
    assert torch.equal(kept._edit.mapper_nz_idx, e.mapper_nz_idx) and torch.equal(kept._edit.mapper_nz_w, e.mapper_nz_w)
    a, b = ["a bowl of soup", "a bowl of hot pea soup"], ["a cat on a table", "a cat on a wooden table"]
    kept, fresh = pkg.p2p.AttentionRefine(a, tok, 5, 0.8, 0.6, device="cpu"), pkg.p2p.AttentionRefine(b, tok, 5, 0.8, 0.6, device="cpu")
    kept._edit = kept.cross_edit()
    kept.retarget(b, tok)
    e = fresh.cross_edit()
    assert torch.equal(kept._edit.mapper_idx, e.mapper_idx) and torch.equal(kept._edit.refine_alpha, e.refine_alpha)
