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
