"""``ModelWrapper`` (robot_flamingo/eval/eval_utils.py:187-490): the CALVIN policy adapter, on the engine.

``step(obs, goal)`` keeps the reference contract -- uint8 camera frames + an instruction string in, a
``(multi_execution, 7)`` float16 numpy action out -- and ``reset()`` clears the episode's LSTM state.
``BatchedModelWrapper`` steps N environments in lock-step through one engine call (slots = environments).
Host-side image preprocessing is the open_clip transform (Pillow bicubic resize to 224, center crop, /255, CLIP mean/std);
``BatchedModelWrapper(device_transform=True)`` and ``DeerEngine.step_u8`` run the same transform, bit-exactly, on the GPU
(SURVEY §8 f1).
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F

from .weights import CLIP_MEAN, CLIP_STD


def clip_image_transform(image_size: int = 224) -> Callable:
    """uint8 HWC numpy / PIL image -> (3, image_size, image_size) f32, like open_clip's image_transform(is_train=False)."""
    mean = torch.tensor(CLIP_MEAN).view(3, 1, 1)
    std = torch.tensor(CLIP_STD).view(3, 1, 1)

    try:
        from PIL import Image
    except ImportError:                       # no Pillow: antialiased torch bicubic (close to, not identical with, the reference)
        Image = None

    def fn_pil(img) -> torch.Tensor:
        # torchvision Resize(size, BICUBIC) -> CenterCrop(size) -> ToTensor -> Normalize, which is what open_clip composes
        im = img if isinstance(img, Image.Image) else Image.fromarray(np.asarray(img, dtype=np.uint8))
        w, h = im.size
        nh, nw = (image_size, int(image_size * w / h)) if h <= w else (int(image_size * h / w), image_size)
        if (nw, nh) != (w, h):
            im = im.resize((nw, nh), Image.BICUBIC)
        t, l = int(round((nh - image_size) / 2.0)), int(round((nw - image_size) / 2.0))
        a = torch.from_numpy(np.asarray(im.convert("RGB"))[t:t + image_size, l:l + image_size].copy()).permute(2, 0, 1).float().div(255)
        return ((a - mean) / std).contiguous()

    def fn(img) -> torch.Tensor:
        if Image is not None:
            return fn_pil(img)
        a = torch.from_numpy(np.asarray(img)).permute(2, 0, 1).float().unsqueeze(0) / 255.0
        h, w = a.shape[-2:]
        s = image_size / min(h, w)
        nh, nw = max(image_size, round(h * s)), max(image_size, round(w * s))
        a = F.interpolate(a, size=(nh, nw), mode="bicubic", align_corners=False, antialias=True).clamp_(0, 1)
        t, l = (nh - image_size) // 2, (nw - image_size) // 2
        a = a[0, :, t:t + image_size, l:l + image_size]
        return ((a - mean) / std).contiguous()
    return fn


def preprocess_image(sample, image_processor) -> torch.Tensor:          # data.py:898-902
    return torch.cat([image_processor(s).unsqueeze(0) for s in sample], dim=0)


def preprocess_text_calvin(sample: Sequence[str], tokenizer):           # data.py:905-919
    tokenizer.padding_side = "right"
    sample = [f"<image>{s.strip()}<|endofchunk|>{tokenizer.eos_token}" for s in sample]
    text = tokenizer(sample, max_length=32, padding="longest", truncation="only_first", return_tensors="pt")
    return text["input_ids"], text["attention_mask"]


def postprocess_action(pose: torch.Tensor, gripper: torch.Tensor, multi_execution: int = 1) -> np.ndarray:
    """eval_utils.py:454-477 (act_step == 1): cat(pose, gripper>0.5)[-1]; gripper -> +-1; repeat; fp16 numpy."""
    action = torch.cat((pose.reshape(-1, 6), (gripper.reshape(-1, 1) > 0.5).float()), dim=1)[-1]
    action[-1] = (action[-1] - 0.5) * 2
    action = torch.stack([action] * multi_execution, dim=0)
    return action.cpu().detach().to(dtype=torch.float16).numpy()


class ModelWrapper:
    def __init__(self, model, tokenizer, image_processor, cast_dtype=None, use_diff=False, history_len=None, future_act_len=-1,
                 amp=False, exit_id=None, early_exit=False, exit_controller=None, multi_execution=1, use_action_ensemble=False):
        assert not use_diff, "deterministic LSTM head only (eval_calvin.py:201)"
        self.model = model
        self.use_action_ensemble = use_action_ensemble
        if use_action_ensemble:
            # the mean over the last two probes' actions is taken inside the exit gate (deer_config.action_ensemble);
            # the reference's get_ensemble_action needs a probe this step, i.e. the dynamic path (value_net.py:92-95)
            assert model.module.engine.action_ensemble, "build the engine with action_ensemble=True (factory: use_action_ensemble)"
            assert early_exit and exit_controller is not None, "use_action_ensemble needs dynamic early exit (value_net.py:93)"
            assert int(model.module.cfg.steps_per_stage) == 1, "action_list is empty on non-boundary steps (value_net.py:285-286,93)"
        self.replan = model.module.replan
        self.decoder_type = model.module.decoder_type
        self.head_type = model.module.head_type
        self.fusion_mode = model.module.fusion_mode
        self.text_process_fn = lambda goals: preprocess_text_calvin(goals, tokenizer)
        self.image_process_fn = lambda imgs: preprocess_image(imgs, image_processor)
        self.amp, self.exit_id = amp, exit_id
        self.dynamic_early_exit, self.exit_controller = early_exit, exit_controller
        self.multi_execution = multi_execution
        self.current_exit_layer = exit_id if isinstance(exit_id, int) else -1
        assert model.module.act_step == 1
        self.device = model.module.engine.device

    def reset(self):                                                    # eval_utils.py:252-277
        self.model.module.clear_all_exit_memory()

    def step(self, obs, goal, get_action=True):                         # eval_utils.py:279-490
        if not get_action:
            return None
        image_x = self.image_process_fn([obs["rgb_obs"]["rgb_static"]]).unsqueeze(1).unsqueeze(1)
        gripper = self.image_process_fn([obs["rgb_obs"]["rgb_gripper"]]).unsqueeze(1).unsqueeze(1)
        text_x, mask = self.text_process_fn([goal])
        assert self.model.module.pad_length == -1
        window_size = self.model.module.set_all_exit_window_size(1)
        out = self.model(vision_x=image_x.to(self.device), lang_x=text_x.to(self.device), attention_mask=mask.to(self.device),
                         vision_gripper=gripper.to(self.device), state_tensor=None, return_feature=True, deterministic=True,
                         exit_id=self.exit_id, dynamic_early_exit=self.dynamic_early_exit, exit_controller=self.exit_controller)
        self.current_exit_layer = out.exit_layer
        if self.use_action_ensemble:                                    # eval_utils.py:460-464: mean of action_list[-2:], gripper > 0.5, +-1
            a = out.action.reshape(-1, 7)[-1]
            action = torch.stack([a] * self.multi_execution, dim=0).cpu().detach().to(dtype=torch.float16).numpy()
        else:
            action = postprocess_action(out.logits[0], out.logits[1], self.multi_execution)
        self.model.module.set_all_exit_window_size(window_size)
        return action

    def on_rollout_step(self, step: int):
        """The per-step bookkeeping of the reference's rollout loop (eval_utils.py:652-663): every ``replan`` steps the policy
        is reset (``refresh == -1``); with ``refresh != -1`` the reference only clears ``lang_encoder.lm_head`` and trims the
        value net's history, neither of which the DeeR exit head reads, so the carried LSTM state is kept.  Then
        ``exit_controller.set_timestep(step)``."""
        m = self.model.module
        if self.replan != -1 and step % self.replan == 0 and m.refresh == -1:
            self.reset()
        if self.exit_controller is not None:
            self.exit_controller.module.set_timestep(step)


class BatchedModelWrapper:
    """N environments in lock-step: one ``deer_step`` per environment step.  ``reset(i)`` clears one env's
    LSTM slot (a new subtask rollout, eval_utils.py:646)."""

    def __init__(self, model, tokenizer, image_processor, exit_controller=None, multi_execution=1, device_transform=False):
        """``device_transform=True``: the raw uint8 camera frames go to the GPU as they are and the open_clip transform
        (Pillow-exact bicubic resize + normalise) runs there (``deer_step_host_u8``); ``image_processor`` is then unused."""
        self.model, self.exit_controller, self.multi_execution = model, exit_controller, multi_execution
        self.device_transform = device_transform
        self.text_process_fn = lambda goals: preprocess_text_calvin(goals, tokenizer)
        self.image_process_fn = lambda imgs: preprocess_image(imgs, image_processor)
        self.device = model.module.engine.device
        self.current_exit_layers: Optional[List[int]] = None

    def reset(self, env_ids: Optional[Sequence[int]] = None):
        self.model.module.engine.reset(env_ids)

    def step(self, obs_list, goals: Sequence[str], env_ids: Optional[Sequence[int]] = None) -> np.ndarray:
        n = len(obs_list)
        if self.device_transform:
            eng = self.model.module.engine
            if self.exit_controller is not None:
                assert self.exit_controller.thresholds is not None, "Please set thresholds before calling forward"     # value_net.py:279
            rgb8 = torch.from_numpy(np.stack([np.asarray(o["rgb_obs"]["rgb_static"], dtype=np.uint8) for o in obs_list]))
            grip8 = torch.from_numpy(np.stack([np.asarray(o["rgb_obs"]["rgb_gripper"], dtype=np.uint8) for o in obs_list]))
            ids, mask = self.text_process_fn(list(goals))
            slots = None if env_ids is None else torch.tensor(list(env_ids), dtype=torch.int32)
            out = eng.step_u8(rgb8, grip8, ids, mask.to(torch.uint8), slots,
                              exit_id=None if self.exit_controller is not None else eng.cfg.n_layers - 1)
            self.current_exit_layers = out["exit_layer"].tolist()
            a = out["action"].clone()
            return torch.stack([a] * self.multi_execution, dim=1).to(torch.float16).numpy()
        rgb = self.image_process_fn([o["rgb_obs"]["rgb_static"] for o in obs_list]).view(n, 1, 1, 3, *([self.model.module.cfg.image_size] * 2))
        grip = self.image_process_fn([o["rgb_obs"]["rgb_gripper"] for o in obs_list]).view(n, 1, 1, 3, *([self.model.module.cfg.image_size] * 2))
        ids, mask = self.text_process_fn(list(goals))
        slots = None if env_ids is None else torch.tensor(list(env_ids), dtype=torch.int32, device=self.device)
        out = self.model(vision_x=rgb.to(self.device), lang_x=ids.to(self.device), attention_mask=mask.to(self.device),
                         vision_gripper=grip.to(self.device), dynamic_early_exit=self.exit_controller is not None,
                         exit_controller=self.exit_controller, slots=slots)
        self.current_exit_layers = out.exit_layers.cpu().tolist()
        a = out.action.cpu()
        return torch.stack([a] * self.multi_execution, dim=1).to(torch.float16).numpy()      # (n, multi_execution, 7)
