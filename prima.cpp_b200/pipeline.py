"""Layer-window pipeline over torch.distributed — prima.cpp's piped ring (src/llama.cpp:3838-3883 this_layer_is_mine /
map_layer_to_local_id, llama_send/recv_tensors :18031-18077) re-targeted to the GPUs of one NVSwitch box.

One process per GPU.  Rank r owns the contiguous layers [bounds[r], bounds[r+1]); rank 0 also owns the token embedding, the
last rank owns output_norm + lm_head.  Per token there is exactly ONE point-to-point hand-off of the hidden state
[n_embd] f32 per stage boundary (NCCL send/recv over NVLink; the reference ships the same f32 payload over ZeroMQ/TCP), and
the sampled token id returns from the last rank to rank 0 (prima returns the result to the master, :18559-18563), which keeps
decoding strictly sequential.  No collective is involved: the path has no reduction step.

The stage object only needs: decode_async(token, pos), and the tensors hidden_in / hidden_out / logits that alias the
engine's device buffers.  The same class drives CPU tensors over gloo in the tests."""
from __future__ import annotations


def layer_windows(n_layer: int, world: int) -> list[int]:
    """Uniform contiguous windows (8 identical B200s need no ILP scheduler, common/common.cpp:860-1594): bounds[r]..bounds[r+1]."""
    if world < 1 or n_layer < world:
        raise ValueError("need 1 <= world <= n_layer")
    return [round(r * n_layer / world) for r in range(world + 1)]


class PipelineRunner:
    def __init__(self, stage, rank: int, world: int, dist=None, token_tensor=None):
        self.stage, self.rank, self.world, self.dist = stage, rank, world, dist
        self.tok = token_tensor   # int64[1] on the stage's device
        if world > 1 and dist is None:
            raise ValueError("world > 1 needs torch.distributed")

    @property
    def first(self) -> bool:
        return self.rank == 0

    @property
    def last(self) -> bool:
        return self.rank == self.world - 1

    def step(self, token: int, pos: int, sample=None) -> int | None:
        """Runs one token through the pipeline.  Returns the sampled token on rank 0 (None elsewhere / when world == 1 and no
        sampler is given).  `sample(logits_tensor) -> int64[1] tensor` runs on the last rank."""
        st, d = self.stage, self.dist
        if self.world == 1:
            st.decode_async(token, pos)
            return None
        if not self.first:
            d.recv(st.hidden_in, src=self.rank - 1)
        st.decode_async(token, pos)
        if not self.last:
            d.send(st.hidden_out, dst=self.rank + 1)
        if self.last:
            self.tok.copy_(sample(st.logits).reshape(1))
            d.send(self.tok, dst=0)
        if self.first:
            d.recv(self.tok, src=self.world - 1)
            return int(self.tok.cpu().item())   # the master must see the token before the next step can start
        return None
