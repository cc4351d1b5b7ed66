"""Host-side logic of the data-parallel layer on CPU: world_size 2, gloo backend (one image per rank,
one weight broadcast at start-up, one final gather, no per-step collective)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from physicedit_b200 import parallel


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(100 + rank)                      # ranks start with DIFFERENT weights
        fused = torch.randn(6, 4)
        m = torch.nn.Sequential(torch.nn.Linear(4, 3), torch.nn.Linear(4, 3))
        m[0].weight.data = fused[0:3]                      # views into one storage, like the engine's fused QKV buffer
        m[1].weight.data = fused[3:6]
        nbytes = parallel.broadcast_weights(m, src=0)
        ref = [torch.zeros_like(fused)]
        if rank == 0:
            ref = [fused.clone()]
        dist.broadcast(ref[0], src=0)
        same = torch.equal(fused, ref[0])
        mine = parallel.shard_indices(5)
        lat = torch.full((len(mine[:2]), 16, 2, 2), float(rank))
        gathered = parallel.gather_latents(lat[:2], dst=0)
        q.put((rank, same, nbytes, mine, None if gathered is None else [float(g.mean()) for g in gathered]))
    finally:
        dist.destroy_process_group()


def test_broadcast_shard_gather_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, same0, n0, mine0, g0), (r1, same1, n1, mine1, g1) = res
    assert same0 and same1                                  # rank 1 now holds rank 0's weights (through the views)
    assert n0 == n1 == (6 * 4 + 3 + 3) * 4                  # shared storage de-duplicated per view, biases separate
    assert mine0 == [0, 2, 4] and mine1 == [1, 3]           # disjoint, exhaustive
    assert g0 == [0.0, 1.0] and g1 is None


def test_shard_indices_cover_everything():
    for n in (0, 1, 7, 8, 9, 64):
        for world in (1, 2, 4, 8):
            parts = [parallel.shard_indices(n, r, world) for r in range(world)]
            flat = sorted(i for p in parts for i in p)
            assert flat == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def test_c_abi_exports_every_declared_symbol():
    """The C-ABI library loads on a CPU-only host and exports every symbol include/pe_b200.h declares."""
    import re
    from physicedit_b200 import native
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lib = native.load_library()
    hdr = open(os.path.join(root, "include", "pe_b200.h")).read()
    declared = set(re.findall(r"\b(pe_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    for sym in declared:
        assert hasattr(lib, sym), f"{sym} declared in pe_b200.h but not exported"
    assert set(native.EXPORTED_SYMBOLS) <= declared
    assert lib.pe_abi_version() == 2
    if not torch.cuda.is_available():
        with pytest.raises(native.NativeUnavailable):      # the product path fails loudly without a GPU: no fallback
            native.Native.get(0)


def test_compat_alias_package_resolves_reference_imports():
    """The import lines of scripts/inference/validate.py:10-12 and scripts/train/train_physicedit.py:1-6 resolve to this package."""
    import sys
    saved = {k: v for k, v in sys.modules.items() if k == "diffsynth" or k.startswith("diffsynth.")}
    try:
        from physicedit_b200 import compat
        compat.install()
        from diffsynth import load_state_dict, ModelConfig                                            # noqa: F401
        from diffsynth.pipelines.qwen_image_physical import QwenImagePhysicPipeline, ModelConfig as MC  # noqa: F401
        from diffsynth.pipelines.flux_image_new import ControlNetInput
        from diffsynth.schedulers.flow_match import FlowMatchScheduler
        assert ControlNetInput().scale == 1.0
        s = FlowMatchScheduler(sigma_min=0, sigma_max=1, extra_one_step=True, exponential_shift=True, exponential_shift_mu=0.8, shift_terminal=0.02)
        assert s.timesteps.max().item() == 1000.0
        cfg = MC(path="/nonexistent/file.safetensors")
        cfg.download_if_necessary()
        assert cfg.path == "/nonexistent/file.safetensors"
    finally:
        for k in [k for k in sys.modules if k == "diffsynth" or k.startswith("diffsynth.")]:
            del sys.modules[k]
        sys.modules.update(saved)


def test_pipeline_host_contract_on_cpu():
    """Host-side behaviours mirrored from the reference that need no GPU."""
    import torch
    from physicedit_b200.lora import GeneralLoRALoader
    from physicedit_b200.pipeline import QwenImagePhysicPipeline, hash_state_dict_keys, DIT_KEY_HASH
    from physicedit_b200.dit import QwenImageDiT
    with torch.device("meta"):
        dit = QwenImageDiT()
    assert hash_state_dict_keys(dit.state_dict()) == DIT_KEY_HASH
    names = dict(dit.named_modules())
    for tgt in ("to_q", "to_k", "to_v", "add_q_proj", "add_k_proj", "add_v_proj", "to_out.0", "to_add_out"):
        assert isinstance(names[f"transformer_blocks.59.attn.{tgt}"], torch.nn.Linear)        # PEFT / LoRA targets are real nn.Linear
    for tgt in ("img_mlp.net.2", "txt_mlp.net.2", "img_mod.1", "txt_mod.1"):
        assert isinstance(names[f"transformer_blocks.0.{tgt}"], torch.nn.Linear)
    loader = GeneralLoRALoader()
    nd = loader.get_name_dict({"diffusion_model.transformer_blocks.3.attn.to_k.lora_B.weight": 0, "transformer_blocks.3.attn.to_q.lora_B.default.weight": 0,
                               "transformer_blocks.3.attn.to_q.lora_A.default.weight": 0})
    assert nd == {"transformer_blocks.3.attn.to_k": ("diffusion_model.transformer_blocks.3.attn.to_k.lora_B.weight", "diffusion_model.transformer_blocks.3.attn.to_k.lora_A.weight"),
                  "transformer_blocks.3.attn.to_q": ("transformer_blocks.3.attn.to_q.lora_B.default.weight", "transformer_blocks.3.attn.to_q.lora_A.default.weight")}
    pipe = QwenImagePhysicPipeline(device="cpu", torch_dtype=torch.bfloat16, build_training_path=False)
    assert pipe.check_resize_height_width(1000, 1030) == (1008, 1040)                          # rounds UP to multiples of 16
    assert pipe.visual_thinking_adapter.t_min == 19.999980926513672 and pipe.visual_thinking_adapter.t_max == 1000.0
    assert pipe.in_iteration_models == ("dit", "blockwise_controlnet", "visual_thinking_adapter")
    n = pipe.generate_noise((1, 16, 4, 4), seed=0)
    assert torch.equal(n, torch.randn((1, 16, 4, 4), generator=torch.Generator("cpu").manual_seed(0)).to(torch.bfloat16))
    with pytest.raises(AssertionError):
        QwenImagePhysicPipeline(device="cpu", dinov2_path=None)                                  # assert dinov2_path is not None (:198)
    # checkpoint key layout of the reference (pipe.-stripped): strict=False load touches exactly the adapter keys
    sd = {"visual_thinking_adapter.head_dino.0.weight": torch.zeros(10752, 3584, dtype=torch.bfloat16)}
    res = pipe.load_state_dict(sd, strict=False)
    assert not res.unexpected_keys


def test_c_abi_rejects_null_handle_without_a_gpu():
    """Error convention of include/pe_b200.h: every entry point returns a negative pe_status (never crashes, never throws)
    when called with a NULL handle -- checked here without any CUDA device."""
    import ctypes
    from physicedit_b200 import native
    lib = native.load_library()
    null = ctypes.c_void_p(None)
    assert lib.pe_destroy(null) == -1
    assert lib.pe_sm_count(null) == -1
    assert lib.pe_last_error(null) == b"null handle"
    assert lib.pe_gemm(null, None, 1, 8, 8, 0, 0, None) == -1
    assert lib.pe_attention_fwd(null, None, None, None, None, 1, 1, 128, ctypes.c_float(1.0), 0, None) == -1
    assert lib.pe_layernorm_modulate(null, None, None, 1, 8, None, None, None) == -1
    assert lib.pe_gemv(null, None, None, None, None, 1, 8, 8, 0, 0, None, None) == -1
    assert lib.pe_timestep_embedding(null, None, None, 1, None) == -1
    assert lib.pe_cfg_euler_step(null, None, None, None, 1, ctypes.c_float(1.0), ctypes.c_float(0.0), None) == -1
    assert lib.pe_check_async_error(null, None, None) == -1
    out = ctypes.c_void_p()
    assert lib.pe_create(None, 0) == -1                              # null out pointer
    if not torch.cuda.is_available():
        assert lib.pe_create(ctypes.byref(out), 0) == -3             # PE_ERR_UNSUPPORTED_DEVICE: no sm_100 device, no fallback
        assert out.value is None


def _cfg_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from physicedit_b200.pipeline import QwenImagePhysicPipeline
        grp, pair, npairs = parallel.make_cfg_pairs()
        calls = []

        class Fake:                                      # stands in for the pipeline: only run_cfg_branches' host logic is under test
            cfg_parallel_group = grp
            cfg_streams = 2

            def model_fn(self, out=None, tag=None, **kw):
                calls.append(tag)
                out.fill_(1.0 if tag == "posi" else -2.0)

        vbuf = torch.zeros(2, 1, 16, 4, 4)
        QwenImagePhysicPipeline.run_cfg_branches(Fake(), {}, {"tag": "posi"}, {"tag": "nega"}, vbuf[0], vbuf[1], None, 0.0)
        q.put((rank, pair, npairs, calls, float(vbuf[0].mean()), float(vbuf[1].mean())))
    finally:
        dist.destroy_process_group()


def test_cfg_parallel_pair_runs_one_branch_per_rank_and_exchanges():
    """SURVEY 8f4 host logic on CPU (gloo, world 2): rank 0 runs only the positive branch, rank 1 only the negative one, and after the
    exchange both ranks hold both predictions."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_cfg_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0] == (0, 0, 1, ["posi"], 1.0, -2.0)
    assert res[1] == (1, 0, 1, ["nega"], 1.0, -2.0)


class _ToyData(torch.utils.data.Dataset):
    load_from_cache = False

    def __init__(self):
        g = torch.Generator().manual_seed(0)
        self.x = torch.randn(4, 6, generator=g)
        self.y = torch.randn(4, 2, generator=g)

    def __len__(self):
        return 4

    def __getitem__(self, i):
        return dict(x=self.x[i], y=self.y[i])


def _toy_module():
    from physicedit_b200.trainers import DiffusionTrainingModule

    class Toy(DiffusionTrainingModule):
        def __init__(self):
            super().__init__()
            torch.manual_seed(7)
            self.pipe = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 2))
            self.pipe[0].requires_grad_(False)                       # a frozen part: must not appear in the checkpoint

        def forward(self, data, inputs=None):
            return torch.nn.functional.mse_loss(self.pipe(data["x"]), data["y"])
    return Toy()


def _train_worker(rank, world, port, q, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from physicedit_b200.trainers import ModelLogger, launch_training_task
        model = _toy_module()
        launch_training_task(_ToyData(), model, ModelLogger(out_dir, remove_prefix_in_ckpt="pipe."), learning_rate=1e-2, weight_decay=0.0, num_workers=0,
                             num_epochs=2, gradient_accumulation_steps=1, find_unused_parameters=False)
        q.put((rank, torch.cat([p.detach().flatten() for p in model.parameters()]).tolist()))      # plain lists: a tensor in the queue needs the sender alive
    finally:
        dist.destroy_process_group()


def test_launch_training_task_world2_matches_single_process_accumulation(tmp_path):
    """launch_training_task (trainers/utils.py:932-977 on plain torch.distributed): 2 ranks x DDP over a DistributedSampler leave identical
    parameters on both ranks, only rank 0 writes the trainable-only checkpoints with the prefix stripped, and the run equals one process that
    accumulates the same two samples per optimizer step (same seeds -> same shuffles)."""
    from safetensors.torch import load_file
    from physicedit_b200.trainers import ModelLogger, launch_training_task
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_train_worker, args=(r, 2, port, q, str(tmp_path / "ddp"))) for r in range(2)]
    for p in procs:
        p.start()
    res = {r: torch.tensor(v) for r, v in (q.get(timeout=180) for _ in procs)}
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert torch.equal(res[0], res[1])
    files = sorted(os.listdir(tmp_path / "ddp"))
    assert files == ["epoch-0.safetensors", "epoch-1.safetensors"]
    ck = load_file(str(tmp_path / "ddp" / "epoch-1.safetensors"))
    assert set(ck) == {"2.weight", "2.bias"}                                           # trainable only, "pipe." stripped
    assert torch.allclose(ck["2.weight"].flatten(), res[0][-12:-2], atol=0) and torch.allclose(ck["2.bias"], res[0][-2:], atol=0)
    # the frozen layer did not move
    torch.manual_seed(7)
    ref0 = torch.nn.Linear(6, 5)
    assert torch.equal(res[0][:30], ref0.weight.detach().flatten())


def _groups_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sp, pair = parallel.make_cfg_sequence_groups()
        # what a step does with them: the half's ranks agree on a value (stand-in for the sequence-parallel forward), then the pair exchanges predictions
        t = torch.tensor([float(rank)])
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=sp)
        vbuf = torch.zeros(2, 3)
        r = dist.get_rank(pair)
        vbuf[r] = t.item() + 100 * r
        parallel.exchange_cfg_predictions(vbuf[0], vbuf[1], r, pair)
        q.put((rank, dist.get_world_size(sp), dist.get_rank(sp), dist.get_world_size(pair), r, t.item(), vbuf[:, 0].tolist()))
    finally:
        dist.destroy_process_group()


def test_cfg_split_sequence_groups_world4():
    """parallel.make_cfg_sequence_groups on 4 gloo ranks: halves {0,1} / {2,3} are the sequence-parallel groups, (0,2) and (1,3) the CFG pairs; rank i of
    the first half holds the positive branch (pair rank 0), its partner the negative one, and after the exchange both hold both predictions."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_groups_worker, args=(r, 4, port, q)) for r in range(4)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, sp_n, sp_r, pair_n, pair_r, half_sum, both in res:
        assert sp_n == 2 and pair_n == 2 and sp_r == rank % 2 and pair_r == rank // 2
        assert half_sum == (1.0 if rank < 2 else 5.0)                                  # 0 + 1 | 2 + 3
        assert both == [1.0, 105.0]                                                    # [positive half's value, negative half's value + 100]
