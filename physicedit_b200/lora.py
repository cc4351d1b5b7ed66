"""LoRA fold at load time: W <- W + alpha * (B @ A), in the pipe dtype on the device.

Mirrors DiffSynth-Studio/diffsynth/lora/__init__.py (GeneralLoRALoader.get_name_dict / load): same key
mapping (`<module>.lora_B[.adapter].weight` -> module name, optional `diffusion_model.` prefix), same
arithmetic (bf16 mm + bf16 add).  The fold writes IN PLACE into the module's weight storage, which for
the attention projections is a view of the engine's fused QKV buffer -- so no re-packing is needed.
There is no per-step LoRA compute.
"""
import torch


class GeneralLoRALoader:
    def __init__(self, device="cpu", torch_dtype=torch.float32):
        self.device = device
        self.torch_dtype = torch_dtype

    def get_name_dict(self, lora_state_dict):
        lora_name_dict = {}
        for key in lora_state_dict:
            if ".lora_B." not in key:
                continue
            parts = key.split(".")
            if len(parts) > parts.index("lora_B") + 2:
                parts.pop(parts.index("lora_B") + 1)          # adapter name ("default")
            parts.pop(parts.index("lora_B"))
            if parts[0] == "diffusion_model":
                parts.pop(0)
            parts.pop(-1)                                      # "weight"
            lora_name_dict[".".join(parts)] = (key, key.replace(".lora_B.", ".lora_A."))
        return lora_name_dict

    @torch.no_grad()
    def load(self, model: torch.nn.Module, state_dict_lora, alpha=1.0):
        updated = 0
        names = self.get_name_dict(state_dict_lora)
        for name, module in model.named_modules():
            if name not in names:
                continue
            up = state_dict_lora[names[name][0]].to(device=self.device, dtype=self.torch_dtype)
            down = state_dict_lora[names[name][1]].to(device=self.device, dtype=self.torch_dtype)
            if up.dim() == 4:
                up, down = up.squeeze(3).squeeze(2), down.squeeze(3).squeeze(2)
                delta = alpha * torch.mm(up, down).unsqueeze(2).unsqueeze(3)
            else:
                delta = alpha * torch.mm(up, down)
            w = module.weight
            w.data.copy_(w.data.to(device=self.device, dtype=self.torch_dtype) + delta)
            updated += 1
        if hasattr(model, "_engine") and model._engine is not None:
            model._engine.invalidate()
        print(f"{updated} tensors are updated by LoRA.")
        return updated


# ------------------------------------------------------------------------------------------------
# un-merged LoRA for training
# ------------------------------------------------------------------------------------------------
class LoRALinear(torch.nn.Module):
    """What `peft.inject_adapter_in_model(LoraConfig(r, lora_alpha, target_modules), model)` puts in place of a targeted nn.Linear
    (DiffSynth-Studio/diffsynth/trainers/utils.py:799-808; peft is an un-pinned dependency that is absent here, so this restates its
    published layer: peft/tuners/lora/layer.py `Linear`): the frozen `base_layer`, `lora_A.default` (kaiming-uniform, a = sqrt(5)) and
    `lora_B.default` (zeros), no dropout at the reference's settings, `y = base(x) + lora_B(lora_A(x)) * (lora_alpha / r)`.  Parameter
    names are PEFT's, so `export_trainable_state_dict` writes the keys `validate.py:44-65` / `GeneralLoRALoader` read back
    (`...to_q.lora_A.default.weight`).  The forward runs on the native GEMMs through physicedit_b200.autograd."""

    def __init__(self, base_layer: torch.nn.Linear, r: int, lora_alpha: float):
        super().__init__()
        import math
        self.base_layer = base_layer
        self.r, self.lora_alpha, self.scaling = r, lora_alpha, lora_alpha / r
        kw = dict(bias=False, device=base_layer.weight.device, dtype=base_layer.weight.dtype)
        self.lora_A = torch.nn.ModuleDict({"default": torch.nn.Linear(base_layer.in_features, r, **kw)})
        self.lora_B = torch.nn.ModuleDict({"default": torch.nn.Linear(r, base_layer.out_features, **kw)})
        torch.nn.init.kaiming_uniform_(self.lora_A["default"].weight, a=math.sqrt(5))
        torch.nn.init.zeros_(self.lora_B["default"].weight)
        base_layer.weight.requires_grad_(False)
        if base_layer.bias is not None:
            base_layer.bias.requires_grad_(False)

    # the attributes the loaders / the engine read on a plain Linear
    @property
    def weight(self):
        return self.base_layer.weight

    @property
    def bias(self):
        return self.base_layer.bias

    @property
    def in_features(self):
        return self.base_layer.in_features

    @property
    def out_features(self):
        return self.base_layer.out_features

    def forward(self, x):
        from . import autograd as ag
        return ag.lora_linear(x, self)

    @torch.no_grad()
    def merge(self):
        """Fold `scaling * B @ A` into the base weight (bf16 mm + bf16 add, the arithmetic of GeneralLoRALoader.load) and return the base layer."""
        w = self.base_layer.weight
        w.data.add_(self.scaling * torch.mm(self.lora_B["default"].weight.to(w.dtype), self.lora_A["default"].weight.to(w.dtype)))
        return self.base_layer


def inject_lora(model: torch.nn.Module, target_modules, r: int, lora_alpha=None):
    """PEFT's target matching (a module is wrapped when its qualified name equals a target or ends with "." + target) over the nn.Linear
    modules of `model`; every other parameter of `model` is frozen, as PEFT's `mark_only_adapters_as_trainable` does."""
    lora_alpha = r if lora_alpha is None else lora_alpha
    targets = list(target_modules)
    hits = [(name, mod) for name, mod in model.named_modules()
            if isinstance(mod, torch.nn.Linear) and any(name == t or name.endswith("." + t) for t in targets)]
    if not hits:
        raise ValueError(f"Target modules {targets} not found in the base model.")
    for p in model.parameters():
        p.requires_grad_(False)
    for name, mod in hits:
        parent_name, _, leaf = name.rpartition(".")
        parent = model.get_submodule(parent_name) if parent_name else model
        setattr(parent, leaf, LoRALinear(mod, r, lora_alpha))
    refresh_lora_flag(model)
    if getattr(model, "_engine", None) is not None:
        model._engine.invalidate()
    return model


def merge_lora(model: torch.nn.Module):
    """Fold every LoRALinear of `model` and put the plain nn.Linear back (the inference engine then runs on the folded weights)."""
    for name, mod in list(model.named_modules()):
        if isinstance(mod, LoRALinear):
            parent_name, _, leaf = name.rpartition(".")
            parent = model.get_submodule(parent_name) if parent_name else model
            setattr(parent, leaf, mod.merge())
    refresh_lora_flag(model)
    if getattr(model, "_engine", None) is not None:
        model._engine.invalidate()
    return model


class HotLoRALinear(torch.nn.Module):
    """The LoRA side of `AutoWrappedLinear` (vram_management/layers.py:98-187), which `pipe.enable_lora_magic()` (:288-305) puts around every
    nn.Linear of the DiT so that `pipe.load_lora(..., hotload=True)` (:265-272) can ATTACH LoRAs without folding them: pairs (A * alpha, B) are
    appended to two lists and the forward is `linear(x) + sum_i x A_i^T B_i^T` (:177-179); `pipe.clear_lora()` empties the lists.  The VRAM
    off-loading half of that class is not needed on a 180 GB device.  With empty lists the module is transparent (the inference engine reads
    `.weight` / `.bias`); with attached LoRAs model_fn takes the un-merged path (physicedit_b200/autograd.py)."""

    def __init__(self, base_layer: torch.nn.Linear, name: str = ""):
        super().__init__()
        self.base_layer = base_layer
        self.name = name
        self.lora_A_weights, self.lora_B_weights = [], []

    @property
    def weight(self):
        return self.base_layer.weight

    @property
    def bias(self):
        return self.base_layer.bias

    @property
    def in_features(self):
        return self.base_layer.in_features

    @property
    def out_features(self):
        return self.base_layer.out_features

    def forward(self, x):
        from . import autograd as ag
        return ag.module_linear(self, x)


def enable_hot_lora(model: torch.nn.Module) -> int:
    """Wrap every plain nn.Linear of `model` (idempotent); returns the number of wrappers now present."""
    for name, mod in list(model.named_modules()):
        if isinstance(mod, torch.nn.Linear) and not name.endswith("base_layer") and "lora_A" not in name and "lora_B" not in name:
            parent_name, _, leaf = name.rpartition(".")
            parent = model.get_submodule(parent_name) if parent_name else model
            if not isinstance(parent, (HotLoRALinear, LoRALinear)):
                setattr(parent, leaf, HotLoRALinear(mod, name))
    return sum(isinstance(m, HotLoRALinear) for m in model.modules())


def hotload_lora(model: torch.nn.Module, lora_state_dict, alpha=1.0) -> int:
    """:265-272: for every wrapped linear `name`, attach (`name.lora_A.default.weight` * alpha, `name.lora_B.default.weight`) when both exist."""
    n = 0
    for name, mod in model.named_modules():
        if isinstance(mod, HotLoRALinear):
            a, b = f"{name}.lora_A.default.weight", f"{name}.lora_B.default.weight"
            if a in lora_state_dict and b in lora_state_dict:
                w = mod.base_layer.weight
                mod.lora_A_weights.append((lora_state_dict[a] * alpha).to(device=w.device, dtype=w.dtype).contiguous())
                mod.lora_B_weights.append(lora_state_dict[b].to(device=w.device, dtype=w.dtype).contiguous())
                n += 1
    refresh_lora_flag(model)
    return n


def clear_hot_lora(model: torch.nn.Module) -> None:
    for mod in model.modules():
        if isinstance(mod, HotLoRALinear):
            mod.lora_A_weights.clear()
            mod.lora_B_weights.clear()
    refresh_lora_flag(model)


def refresh_lora_flag(model: torch.nn.Module) -> None:
    """`_lora_injected` tells model_fn to run the un-merged path: PEFT-style factors present, or a hot LoRA attached."""
    model._lora_injected = any(isinstance(m, LoRALinear) or (isinstance(m, HotLoRALinear) and m.lora_A_weights) for m in model.modules())
