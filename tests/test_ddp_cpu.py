"""CPU, world_size 2 over gloo: the N>1 host logic of the hot path -- FlatDDP bucketing of the flat gradient buffer
(trainer.py:76-88 / distributed/ddp_model_wrapper.py semantics: SUM then 1/world folded into the optimizer, no_sync on
non-final micro-steps, rank-0 weight and BatchNorm-buffer broadcast)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, q) -> None:
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    try:
        from liteasr_b200.distributed.flat_ddp import FlatDDP
        from liteasr_b200.distributed.utils import distributed_init, get_rank, get_world_size
        from liteasr_b200.models.u2 import U2, U2Config
        from liteasr_b200.store import ParamStore
        distributed_init("gloo")
        assert get_world_size() == world and get_rank() == rank
        torch.manual_seed(100 + rank)  # different initial weights per rank: the constructor must broadcast rank 0's
        m = U2(U2Config(input_dim=80, vocab_size=50, enc_layers=2, dec_layers=1, enc_dim=64, dec_dim=64, enc_attn_heads=1,
                        dec_attn_heads=1, enc_ff_dim=96, dec_ff_dim=96))
        st = ParamStore(m, torch.device("cpu"), "fp32")
        ddp = FlatDDP(m, st, bucket_bytes=64 << 10)
        w0 = st.flat.clone()
        ws = [torch.empty_like(w0) for _ in range(world)]
        dist.all_gather(ws, w0)
        assert all(torch.equal(w, ws[0]) for w in ws), "weights not broadcast from rank 0"
        # BatchNorm running stats live in one flat buffer and follow rank 0
        bn = [x for x in m.modules() if isinstance(x, torch.nn.BatchNorm1d)][0]
        bn.running_mean.fill_(float(rank + 1))
        ddp.broadcast_buffers()
        assert float(bn.running_mean[0]) == 1.0

        def fake_backward():
            st.gflat.copy_(torch.arange(st.numel, dtype=torch.float32) * 1e-3 + (rank + 1))
            ddp.begin_backward()
            # the engine announces ranges in backward order: CTC head, decoder, encoder layers last -> first, front end
            for pfx in ("ctc.", "decoder.", "encoder.after_norm.", "encoder.enc_layers.1.", "encoder.enc_layers.0.", "encoder.embed."):
                st.grad_ready_hook(*st.range_of(pfx))
            return ddp.finish_backward()

        mult = fake_backward()
        assert mult == 1.0 / world
        want = torch.arange(st.numel, dtype=torch.float32) * 1e-3 * world + sum(range(1, world + 1))
        assert torch.allclose(st.gflat, want), "all-reduce must cover every element exactly once"
        cover = sorted(ddp.launched)
        assert cover[0][0] == 0 and cover[-1][1] == st.numel and all(a[1] == b[0] for a, b in zip(cover, cover[1:]))
        assert len(cover) >= 2, "expected several buckets"
        # no_sync micro-step: nothing is reduced, multiplier 1
        ddp.sync_grads = False
        mult = fake_backward()
        assert mult == 1.0 and not ddp.launched
        assert torch.allclose(st.gflat, torch.arange(st.numel, dtype=torch.float32) * 1e-3 + (rank + 1))
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        import traceback
        q.put((rank, "".join(traceback.format_exception(e))))


def test_flat_ddp_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, msg in res:
        assert msg == "ok", f"rank {rank}: {msg}"
