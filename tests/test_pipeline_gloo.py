"""CPU: the N>1 host logic (layer windows + hidden-state hand-off + token ring) over gloo, world_size 2 and 3, against a
single-process run of the same toy layers (bit-identical, like the N-rank == 1-rank requirement of SURVEY §8e)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


class ToyStage:
    """Stands in for the engine: layer l multiplies by a fixed matrix and adds the position; exact in fp32 order."""
    E, V, L = 16, 11, 6

    def __init__(self, l0, l1, first, last):
        g = torch.Generator().manual_seed(0)
        self.W = [torch.randn(self.E, self.E, generator=g) / 4 for _ in range(self.L)]
        self.emb = torch.randn(self.V, self.E, generator=g)
        self.head = torch.randn(self.V, self.E, generator=g)
        self.l0, self.l1, self.first, self.last = l0, l1, first, last
        self.hidden_in, self.hidden_out, self.logits = torch.zeros(self.E), torch.zeros(self.E), torch.zeros(self.V)

    def decode_async(self, token, pos):
        x = self.emb[token].clone() if self.first else self.hidden_in.clone()
        for l in range(self.l0, self.l1):
            x = torch.tanh(self.W[l] @ x) + 0.01 * pos
        self.hidden_out.copy_(x)
        if self.last:
            self.logits.copy_(self.head @ x)


def _worker(rank, world, port, q):
    import sys
    from pathlib import Path
    sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
    import pkgload
    import torch.distributed as dist
    pkg = pkgload.load()
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    b = pkg.layer_windows(ToyStage.L, world)
    st = ToyStage(b[rank], b[rank + 1], rank == 0, rank == world - 1)
    run = pkg.PipelineRunner(st, rank, world, dist, torch.zeros(1, dtype=torch.int64))
    tok, toks, outs = 3, [], []
    for pos in range(5):
        nxt = run.step(tok, pos, sample=lambda lg: torch.argmax(lg))
        if rank == world - 1:
            outs.append(st.logits.clone())
        # every rank needs the same next token to stay in lock-step: rank 0 broadcasts it (as llama_send_meta does, :17870)
        t = torch.tensor([nxt if rank == 0 else 0], dtype=torch.int64)
        dist.broadcast(t, src=0)
        tok = int(t.item())
        toks.append(tok)
    if rank == world - 1:
        q.put((toks, torch.stack(outs)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_pipeline_matches_single_process(world):
    single = ToyStage(0, ToyStage.L, True, True)
    tok, want_toks, want = 3, [], []
    for pos in range(5):
        single.decode_async(tok, pos)
        want.append(single.logits.clone())
        tok = int(torch.argmax(single.logits))
        want_toks.append(tok)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in procs]
    toks, outs = q.get(timeout=120)
    [p.join(60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    assert toks == want_toks
    assert torch.equal(outs, torch.stack(want))


def test_layer_windows(pkg):
    assert pkg.layer_windows(80, 1) == [0, 80]
    assert pkg.layer_windows(80, 8) == [0, 10, 20, 30, 40, 50, 60, 70, 80]
    b = pkg.layer_windows(32, 3)
    assert b[0] == 0 and b[-1] == 32 and all(b[i] < b[i + 1] for i in range(3))
    with pytest.raises(ValueError):
        pkg.layer_windows(2, 3)


# ---------------------------------------------------------------------------------------------------------------------
# RingRunner: one sequence per stage in flight (the mode bench.py reports at N > 1)
class ToyRingStage(ToyStage):
    def __init__(self, l0, l1, first, last, n_seq):
        super().__init__(l0, l1, first, last)
        self.tok = [torch.zeros(1, dtype=torch.int32) for _ in range(n_seq)]      # stage 0: token of the slot
        self.pos = [0] * n_seq
        self.smp = [torch.zeros(1, dtype=torch.int32) for _ in range(n_seq)]      # last stage: greedy sample of the slot
        self.history = [[] for _ in range(n_seq)]

    def token_in(self, s):
        return self.tok[s]

    def token_out(self, s):
        return self.smp[s]

    def begin(self, s, token, pos):
        self.tok[s][0] = token
        self.pos[s] = pos

    def run(self, s):
        self.decode_async(int(self.tok[s][0]), self.pos[s])
        self.pos[s] += 1
        if self.last:
            self.smp[s][0] = int(torch.argmax(self.logits))
            self.history[s].append(int(self.smp[s][0]))


def _ring_worker(rank, world, port, q, rounds):
    import sys
    from pathlib import Path
    sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
    import pkgload
    import torch.distributed as dist
    pkg = pkgload.load()
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    b = pkg.layer_windows(ToyStage.L, world)
    st = ToyRingStage(b[rank], b[rank + 1], rank == 0, rank == world - 1, world)
    rr = pkg.RingRunner(st, rank, world, dist)
    seeds = [(3 + s, 2 * s) for s in range(world)]
    rr.slots(world * rounds + world - 1, first_tokens=seeds)     # + world - 1: let the last stage finish the last round
    if rank == world - 1:
        q.put(st.history)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_ring_of_sequences_matches_sequential_decoding(world):
    rounds = 4
    want = []
    for s in range(world):
        single = ToyStage(0, ToyStage.L, True, True)
        tok, pos, hist = 3 + s, 2 * s, []
        for _ in range(rounds):
            single.decode_async(tok, pos)
            tok = int(torch.argmax(single.logits)); pos += 1
            hist.append(tok)
        want.append(hist)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_ring_worker, args=(r, world, port, q, rounds)) for r in range(world)]
    [p.start() for p in procs]
    got = q.get(timeout=120)
    [p.join(60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    assert [h[:rounds] for h in got] == want


# ---- prompt processing through the pipeline in micro-batches (PrefillPipeline; the engine side is pb200_prefill_stage) ----
class ToyPrefillStage:
    """A stage whose 'layers' are an affine map per layer that also depends on the token position, so that a wrong micro-batch order,
    a wrong pos0 or a stale receive buffer changes the result."""
    E, UB = 8, 4

    def __init__(self, l0, l1, first):
        self.l0, self.l1, self.first = l0, l1, first
        self.hidden_buf = torch.zeros((self.UB, self.E), dtype=torch.float64)
        self.calls = []

    def prefill_stage(self, tokens, hidden_in, n, pos0):
        self.calls.append((n, pos0))
        if self.first:
            x = torch.stack([torch.arange(self.E, dtype=torch.float64) * 0.01 + float(t) for t in tokens])
        else:
            x = hidden_in.clone()
        pos = torch.arange(pos0, pos0 + n, dtype=torch.float64)[:, None]
        for il in range(self.l0, self.l1):
            x = x * (1.0 + 0.05 * (il + 1)) + 0.001 * (il + 1) * pos + torch.roll(x, 1, dims=1) * 0.1
        return x


def _prefill_worker(rank, world, port, q, n_tokens):
    import sys
    from pathlib import Path
    sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
    import pkgload
    import torch.distributed as dist
    pkg = pkgload.load()
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    b = pkg.layer_windows(ToyStage.L, world)
    st = ToyPrefillStage(b[rank], b[rank + 1], rank == 0)
    pipe = pkg.PrefillPipeline(st, rank, world, dist, ubatch=ToyPrefillStage.UB)
    toks = [(7 * i + 3) % 11 for i in range(n_tokens)]
    outs = []
    orig = st.prefill_stage

    def spy(tokens, hidden_in, n, pos0):
        y = orig(tokens, hidden_in, n, pos0)
        outs.append(y.clone())
        return y
    st.prefill_stage = spy
    pipe.run(toks, n_tokens, pos0=2)
    if rank == world - 1:
        q.put((torch.cat(outs).tolist(), st.calls))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n_tokens", [(2, 10), (3, 9), (2, 4)])
def test_prefill_pipeline_micro_batches_match_single_stage(world, n_tokens):
    """N stages x ceil(n / ubatch) micro-batches (the last one ragged) over gloo == one stage over the whole prompt."""
    single = ToyPrefillStage(0, ToyStage.L, True)
    toks = [(7 * i + 3) % 11 for i in range(n_tokens)]
    want = single.prefill_stage(toks, None, n_tokens, 2).tolist()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_prefill_worker, args=(r, world, port, q, n_tokens)) for r in range(world)]
    [p.start() for p in procs]
    got, calls = q.get(timeout=120)
    [p.join(60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    assert np.allclose(np.array(got), np.array(want), rtol=0, atol=1e-12)
    ub = ToyPrefillStage.UB
    assert calls == [(min(ub, n_tokens - j), 2 + j) for j in range(0, n_tokens, ub)]
