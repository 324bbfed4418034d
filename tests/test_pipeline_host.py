"""Stage runner on CPU/gloo (BASELINE config 1): argument validation, single-process run, and the
world-size invariance property — the final latent is bit-identical (same SHA-256) whether the schedule
runs on 1, 2 or 4 ranks — including the uneven 25-step split."""
import hashlib
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from vdpp_b200.distributed import finalize_distributed, init_distributed, resolve_backend
from vdpp_b200.models import DummyUNet
from vdpp_b200.pipeline import (LatentSpec, PipelineConfig, PipelineStage, run_pipeline_latents, run_single_latent)

SHAPE = (1, 4, 3, 8, 8)


def _spec():
    return LatentSpec(shape=torch.Size(SHAPE), dtype=torch.float32, device=torch.device("cpu"))


def _model():
    torch.manual_seed(1234)          # the reference does not seed (Q4); every rank must, to agree
    return DummyUNet(channels=4).eval()


def _supplier(i):
    g = torch.Generator().manual_seed(42 + i)
    return torch.randn(SHAPE, generator=g)


def _sha(t):
    return hashlib.sha256(t.contiguous().numpy().tobytes()).hexdigest()


def test_backend_resolution(monkeypatch):
    monkeypatch.delenv("PIPELINE_BACKEND", raising=False)
    assert resolve_backend(None, simulator=True) == "gloo"
    assert resolve_backend(None) == "nccl"
    assert resolve_backend("GLOO") == "gloo"
    monkeypatch.setenv("PIPELINE_BACKEND", "gloo")
    assert resolve_backend(None) == "gloo"
    assert resolve_backend("nccl") == "nccl"          # explicit argument wins over the environment
    with pytest.raises(ValueError):
        resolve_backend("mpi")
    monkeypatch.setenv("PIPELINE_BACKEND", "ucc")
    with pytest.raises(ValueError):
        resolve_backend(None)


def test_config_and_argument_validation():
    with pytest.raises(ValueError):
        PipelineConfig(total_steps=4, world_size=1, rank=0, timesteps=[0, 1, 2], latent_spec=_spec())
    cfg = PipelineConfig(total_steps=4, world_size=1, rank=0, timesteps=[3, 2, 1, 0], latent_spec=_spec())
    st = PipelineStage(_model(), cfg)
    assert (st.step_range.start, st.step_range.end) == (0, 4)
    with pytest.raises(ValueError):
        st.run(None)                                   # rank 0 needs an input
    with pytest.raises(ValueError):
        st.run_many(0, input_supplier=_supplier)
    with pytest.raises(ValueError):
        st.run_many(2)                                 # rank 0 needs a supplier
    with pytest.raises(ValueError):                    # reference rule: 25 steps on 4 ranks is an error ...
        PipelineStage(_model(), PipelineConfig(25, 4, 0, list(range(25)), _spec()))
    ok = PipelineStage(_model(), PipelineConfig(25, 4, 0, list(range(25)), _spec(), allow_uneven=True))
    assert ok.step_range.count == 7                    # ... unless the uneven extension is requested
    mid = PipelineStage(_model(), PipelineConfig(4, 2, 1, [3, 2, 1, 0], _spec()))
    with pytest.raises(ValueError):
        mid._process_single_latent(torch.zeros(SHAPE), 0)   # non-zero ranks must not be handed a latent


def test_single_rank_passes_timestep_values_in_order():
    seen = []

    class Probe(torch.nn.Module):
        def forward(self, latent, step):
            seen.append(step)
            return latent + 1

    ts = [9, 7, 5, 3]
    out = run_single_latent(Probe(), total_steps=4, timesteps=ts, world_size=1, rank=0, latent_spec=_spec(),
                            input_latent=torch.zeros(SHAPE))
    assert seen == ts and torch.equal(out, torch.full(SHAPE, 4.0))       # values, not indices (Q2)
    outs = run_pipeline_latents(Probe(), total_steps=4, timesteps=ts, world_size=1, rank=0, latent_spec=_spec(),
                                num_samples=3, input_supplier=lambda i: torch.full(SHAPE, float(i)))
    assert [o.flatten()[0].item() for o in outs] == [4.0, 5.0, 6.0]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total_steps, n_samples, uneven, use_run_many, q):
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    torch.set_num_threads(1)
    init_distributed(backend="gloo", rank=rank, world_size=world, init_method=f"tcp://127.0.0.1:{port}")
    model = _model()
    ts = list(reversed(range(total_steps)))            # simulator convention (simulator.py:77-79)
    cfg = PipelineConfig(total_steps, world, rank, ts, _spec(), allow_uneven=uneven)
    stage = PipelineStage(model, cfg)
    with torch.no_grad():
        if use_run_many:
            outs = stage.run_many(n_samples, input_supplier=_supplier if rank == 0 else None)
        else:                                           # the reference benchmark's call pattern (benchmark.py:229-235)
            outs = []
            for i in range(n_samples):
                o = stage._process_single_latent(_supplier(i) if rank == 0 else None, sample_idx=i)
                if o is not None:
                    outs.append(o)
            outs = outs or None
    if rank == world - 1:
        q.put([_sha(o) for o in outs])
    else:
        assert outs is None                             # Q10: only the last rank returns latents
    dist.barrier()
    finalize_distributed()


def _ring_worker(rank, world, port, total_steps, n_samples, q):
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    torch.set_num_threads(1)
    init_distributed(backend="gloo", rank=rank, world_size=world, init_method=f"tcp://127.0.0.1:{port}")
    ts = list(reversed(range(total_steps)))
    stage = PipelineStage(_model(), PipelineConfig(total_steps, world, rank, ts, _spec(), allow_uneven=True))
    with torch.no_grad():
        outs = stage.run_many_ring(n_samples, input_supplier=_supplier)
    q.put([(v, _sha(o)) for v, o in outs])
    dist.barrier()
    finalize_distributed()


def _run_ring(world, total_steps, n_samples):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_ring_worker, args=(r, world, port, total_steps, n_samples, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = {}
    for _ in range(world):
        got.update(dict(q.get(timeout=180)))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return [got[v] for v in range(n_samples)]


def _run_world(world, total_steps, n_samples=3, uneven=False, use_run_many=True):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total_steps, n_samples, uneven, use_run_many, q))
             for r in range(world)]
    for p in procs:
        p.start()
    res = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return res


def _single_process(total_steps, n_samples=3):
    model = _model()
    ts = list(reversed(range(total_steps)))
    out = []
    with torch.no_grad():
        for i in range(n_samples):
            x = _supplier(i)
            for t in ts:
                x = model(x, t)
            out.append(_sha(x))
    return out


VERIFY_CASES = ("same", "shape", "dtype", "split")


def _verify_worker(rank, world, port, q):
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    torch.set_num_threads(1)
    init_distributed(backend="gloo", rank=rank, world_size=world, init_method=f"tcp://127.0.0.1:{port}")
    for case in VERIFY_CASES:
        spec = _spec()
        total, uneven = 28, False
        if case == "shape" and rank == 1:
            spec = LatentSpec(shape=torch.Size(list(spec.shape[:-1]) + [spec.shape[-1] + 1]), dtype=spec.dtype, device=spec.device)
        if case == "dtype" and rank == 1:
            spec = LatentSpec(shape=spec.shape, dtype=torch.float64, device=spec.device)
        if case == "split":            # rank 0 splits 25 steps unevenly (13 + 12), rank 1 believes in 24 (12 + 12)
            total, uneven = (25, True) if rank == 0 else (24, False)
        stage = PipelineStage(_model(), PipelineConfig(total, world, rank, list(range(total)), spec, allow_uneven=uneven))
        try:
            stage.verify_peers()
            q.put((case, rank, "ok"))
        except RuntimeError as e:
            q.put((case, rank, str(e)))
    dist.barrier()
    finalize_distributed()


@pytest.mark.timeout(300)
def test_verify_peers_names_the_disagreement_on_every_rank():
    """SURVEY section 5 (failure detection): stages that disagree on the latent or on the split raise at once, on every rank,
    instead of hanging in recv until the process-group timeout as the reference does."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_verify_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = {}
    for _ in range(2 * len(VERIFY_CASES)):
        case, rank, msg = q.get(timeout=180)
        got[(case, rank)] = msg
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got[("same", 0)] == got[("same", 1)] == "ok"
    for case in VERIFY_CASES[1:]:
        a, b = got[(case, 0)], got[(case, 1)]
        assert a == b and a.startswith("pipeline stages disagree: ") and "rank 1" in a, (case, a, b)
    # one process: nothing to compare
    PipelineStage(_model(), PipelineConfig(4, 1, 0, [3, 2, 1, 0], _spec())).verify_peers()


def _negotiate_worker(rank, world, port, total_steps, n_samples, q):
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    torch.set_num_threads(1)
    init_distributed(backend="gloo", rank=rank, world_size=world, init_method=f"tcp://127.0.0.1:{port}")
    ts = list(reversed(range(total_steps)))
    stage = PipelineStage(_model(), PipelineConfig(total_steps, world, rank, ts, _spec()), transport="peer")
    note = stage.negotiate_transport()        # CPU tensors: the peer-mapped slots cannot exist -> every rank drops to send / recv
    assert note is not None and stage.transport == "nccl" and stage._peer is None
    assert stage.negotiate_transport() is None            # nothing left to negotiate
    with torch.no_grad():
        outs = stage.run_many(n_samples, input_supplier=_supplier if rank == 0 else None)
    if rank == world - 1:
        q.put([_sha(o) for o in outs])
    dist.barrier()
    finalize_distributed()


@pytest.mark.timeout(300)
def test_transport_negotiation_falls_back_on_every_rank():
    """``transport="peer"`` where the peer-mapped handoff cannot be set up (CPU / gloo here; a box without symmetric memory
    in production): ``negotiate_transport`` is collective, all ranks agree on send / recv, and the results are unchanged."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_negotiate_worker, args=(r, 2, port, 28, 2, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == _single_process(28, 2)
    # one process: nothing to negotiate, whatever the transport
    solo = PipelineStage(_model(), PipelineConfig(4, 1, 0, [3, 2, 1, 0], _spec()), transport="peer")
    assert solo.negotiate_transport() is None and solo.transport == "peer"


@pytest.mark.timeout(300)
def test_world_size_invariance_even_split():
    want = _single_process(28)
    assert _run_world(2, 28) == want
    assert _run_world(4, 28) == want
    assert _run_world(4, 28, use_run_many=False) == want


@pytest.mark.timeout(300)
def test_world_size_invariance_uneven_split():
    want = _single_process(25)
    assert _run_world(2, 25, uneven=True) == want
    assert _run_world(4, 25, uneven=True) == want          # stages of 7, 6, 6, 6 steps (BASELINE config 1)


@pytest.mark.timeout(300)
def test_ring_placement_matches_single_process():
    """Rotating stage placement: every video still sees the steps in order; 25 steps on 4 ranks (7,6,6,6),
    5 videos (a partial last round), bit-identical to running each video alone."""
    assert _run_ring(4, 25, 5) == _single_process(25, 5)
    assert _run_ring(2, 28, 4) == _single_process(28, 4)


def test_stage_runner_hands_whole_slice_to_forward_steps():
    """A model that offers ``forward_steps`` and sets ``use_stage_graph`` gets its stage's slice of the schedule in ONE
    call (the whole-stage CUDA graph of StableVideoUNet); without the flag the reference's per-step loop runs."""
    import torch
    from vdpp_b200.pipeline import LatentSpec, PipelineConfig, PipelineStage

    class M(torch.nn.Module):
        use_stage_graph = False

        def __init__(self):
            super().__init__()
            self.calls = []

        def forward(self, x, step):
            self.calls.append(("step", step))
            return x + step

        def forward_steps(self, x, steps):
            self.calls.append(("slice", tuple(steps)))
            for s in steps:
                x = x + s
            return x

    spec = LatentSpec(shape=torch.Size((1, 2)), dtype=torch.float32, device=torch.device("cpu"))
    ts = [9, 7, 5, 3]
    outs = []
    for flag in (False, True):
        m = M()
        m.use_stage_graph = flag
        st = PipelineStage(m, PipelineConfig(total_steps=4, world_size=1, rank=0, timesteps=ts, latent_spec=spec))
        outs.append(st.run(torch.zeros(1, 2)))
        assert m.calls == ([("slice", (9, 7, 5, 3))] if flag else [("step", t) for t in ts])
    assert torch.equal(outs[0], outs[1])
