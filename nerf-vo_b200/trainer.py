"""Mapping-step engine: one call = zero grads, forward, losses, backward, (gradient all-reduce,) fused Adam — what
`Nerfstudio.train()` drives through Trainer.train_iteration (nerf_vo/mapping/nerfstudio.py:142-173,
NS/engine/trainer.py:455-494, NS/pipelines/base_pipeline.py:291-304), restated for the nvo_b200 kernels:

  * every parameter lives in ONE flat fp32 buffer, gradients in another; backward kernels scatter straight into the flat
    gradient (no per-tensor .grad, no AccumulateGrad pass), a single fused Adam launch updates everything;
  * the whole step is captured in CUDA graphs (at 4096 rays the step is launch-bound) and replayed;
  * data parallel: rays are sharded across ranks, parameters replicated; the gradient exchange is fused with the optimizer
    (`exchange="fused"`, csrc/exchange.cu: reduce-scatter by NVLink peer loads -> Adam on the rank's slice -> all-gather by
    peer stores, one kernel, the step stays a single CUDA graph) or, as the library arm, ONE NCCL sum-all-reduce of the flat
    gradient followed by the replicated Adam (`exchange="nccl"`); both give DDP's mean-gradient semantics
    (NS/pipelines/base_pipeline.py:281-283);
  * two optimizer parameter groups like the reference's ("fields", "proposal_networks": NS/models/nerfacto.py:244-249, one Adam each,
    NS/engine/optimizers.py:138-150), each with its own step counter.  On 1-2 GPUs both groups are stepped after the backward (fused arm:
    in ONE exchange launch); from 4 ranks on the fields group is exchanged as soon as the main hash-table scatter has landed — on a
    high-priority stream, next to the proposal networks' backward, which is then ordered behind the field's backward chain — and the
    proposal group follows.  `proposal_update="reference"` follows ProposalNetworkSampler's update schedule
    (NS/model_components/ray_samplers.py:596-610): on steps where the proposal networks receive no gradient their backward is not run
    and their Adam group is not stepped (torch skips parameters whose .grad is None); "always" updates them every step.
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional

import torch
import torch.distributed as dist

from . import ops, sharding
from .field_components import repack
from .model import ExtendedNerfactoModel
from .rays import RayBundle


class MappingTrainer:
    def __init__(self, model: ExtendedNerfactoModel, num_rays: int, lr: float = 1e-2, eps: float = 1e-15, betas=(0.9, 0.999),
                 use_cuda_graph: bool = True, with_normals: bool = True, device: Optional[torch.device] = None, exchange: str = "fused",
                 datamanager=None, proposal_update: str = "always", external_draws: bool = False, camera_opt_lr: float = 1e-4,
                 camera_opt_lr_final: float = 1e-5, max_num_iterations: int = 8192, defer_fields_update: bool = False):
        """camera_opt_lr / camera_opt_lr_final / max_num_iterations: the "camera_opt" group (Adam, ExponentialDecayScheduler over the mapping
        iterations, nerf_vo/mapping/nerfstudio.py:93-100), trained when the model's camera optimizer is on (config.camera_optimizer_mode)."""
        self.model = model
        # defer_fields_update: the "fields" group's optimizer (at N > 1: its NVLink exchange) of step k runs at the START of step k+1, next to that
        # step's proposal sampling — which reads the proposal networks only — and is joined before the field forward.  Same arithmetic, same
        # order of updates; what changes is when the fields parameters become current: after train_step() they lag by the last update until
        # the next step or flush() (checkpointing / rendering / reading parameters go through flush()).
        self._defer_requested = bool(defer_fields_update)
        self._pending_fields = False
        # optional: a DynamicDataManager (data.py). The step then starts with the fused prologue kernel (pixel sampling + gather + ray
        # generation, drawn on the device like the reference's torch.rand) instead of reading the static input buffers.
        self.datamanager = datamanager
        # external_draws: the step's random numbers (pixel draws u [B,3], the three stratified jitters) come from the caller through
        # set_draws_packed() instead of being drawn on the device inside the step — a reproducible run, or a host that owns the RNG
        self.external_draws = bool(external_draws)
        self.device = device or next(model.parameters()).device
        self.B = int(num_rays)
        self.lr, self.eps, self.betas = lr, eps, betas
        self.world_size = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        if exchange not in ("fused", "nccl"):
            raise ValueError(f"exchange must be 'fused' (peer-memory reduce-scatter + Adam + all-gather kernel) or 'nccl', got {exchange!r}")
        self.exchange = exchange if self.world_size > 1 else "local"
        self.peer = None
        self._peer_cam = None
        self.use_cuda_graph = use_cuda_graph
        self.with_normals = with_normals
        if proposal_update not in ("always", "reference"):
            raise ValueError(f"proposal_update must be 'always' or 'reference', got {proposal_update!r}")
        self.proposal_update = proposal_update
        self.iteration = 0  # host-side step number (drives the proposal update schedule and the proposal-weight annealing)
        self._ssu = 0       # ProposalNetworkSampler._steps_since_update as the reference's callbacks would leave it
        model.train()
        # ---- flat parameter / gradient / optimizer-state buffers ---------------------------------------------
        self.params = [p for p in model.parameters() if p.requires_grad]
        uniq, seen = [], set()
        for p in self.params:
            if id(p) not in seen:
                seen.add(id(p))
                uniq.append(p)
        self.params = uniq
        # 16-byte alignment of every tensor start keeps float4 / float2 accesses legal: pad each to a multiple of 4 floats
        sizes = [(p.numel() + 3) // 4 * 4 for p in self.params]
        total = sum(sizes)
        # parameter groups = contiguous ranges of the flat buffer: "fields" first, "proposal_networks" behind it (registration order)
        names = {id(p): n for n, p in model.named_parameters()}
        # the camera optimizer's pose deltas ("camera_opt": own learning rate and schedule) sit behind both groups, registered last
        is_cam = [names[id(p)].startswith("camera_optimizer.") for p in self.params]
        n_cam = sum(sz for sz, cm in zip(sizes, is_cam) if cm)
        if n_cam and not all(is_cam[is_cam.index(True):]):
            raise RuntimeError("internal: camera optimizer parameters must be registered last")
        is_prop = [names[id(p)].startswith("proposal_networks.") for p in self.params]
        n_fields = sum(sz for sz, pr, cm in zip(sizes, is_prop, is_cam) if not pr and not cm)
        n_main = total - n_cam
        contiguous = all(not pr for pr in is_prop[:is_prop.index(True)]) and all(pr or cm for pr, cm in zip(is_prop[is_prop.index(True):], is_cam[is_prop.index(True):])) if any(is_prop) else True
        if any(is_prop) and contiguous and 0 < n_fields < n_main:
            self.groups = [("fields", 0, n_fields), ("proposal_networks", n_fields, n_main - n_fields)]
        else:
            self.groups = [("fields", 0, n_main)]
        self.cam_group = ("camera_opt", n_main, n_cam) if n_cam else None
        self.cam_lr, self.cam_lr_final, self.max_num_iterations = float(camera_opt_lr), float(camera_opt_lr_final), int(max_num_iterations)
        if self.exchange == "fused":
            # parameters and gradients live in peer-mapped (CUDA IPC) allocations; Adam moments only for the slices this rank owns
            from .peer import PeerBuffers

            self.peer = PeerBuffers(total, self.device)
            self.flat, self.grad = self.peer.params, self.peer.grads
            self._peer_groups = [self.peer.add_group(off, n) for _, off, n in self.groups]
            self._peer_cam = self.peer.add_group(self.cam_group[1], self.cam_group[2]) if self.cam_group else None
            self.exp_avg = self.exp_avg_sq = None
        else:
            self.flat = torch.zeros(total, dtype=torch.float32, device=self.device)
            self.grad = torch.zeros_like(self.flat)
            self.exp_avg = torch.zeros_like(self.flat)
            self.exp_avg_sq = torch.zeros_like(self.flat)
        # one device step counter per group (torch keeps `step` per parameter; a group that got no gradient is not stepped)
        self.step_counts = [torch.zeros(1, dtype=torch.int32, device=self.device) for _ in self.groups]
        self.step_count = self.step_counts[0]
        self.cam_step = torch.zeros(1, dtype=torch.int32, device=self.device) if self.cam_group else None
        # d loss / d (ray origins, ray directions) of the step in flight (ops.ray_grad_sink) and the pose backward's per-camera scratch
        self._ray_grads = torch.zeros((2, self.B, 3), dtype=torch.float32, device=self.device) if self.cam_group else None
        if self.cam_group and datamanager is not None:
            datamanager.camera_optimizer = model.camera_optimizer  # the step prologue applies the pose correction itself
        self._opt_stream: Optional[torch.cuda.Stream] = None
        self.defer_fields = self._defer_requested and len(self.groups) == 2 and self.exchange != "nccl" and self.device.type == "cuda"
        self._fields_done = False
        self._in_backward = False
        off = 0
        self._views = []
        for p, n in zip(self.params, sizes):
            self.flat[off:off + p.numel()].copy_(p.data.reshape(-1))
            p.data = self.flat[off:off + p.numel()].view(p.shape)
            self._views.append((off, p.numel()))
            off += n
        if self.world_size > 1:
            # DistributedDataParallel broadcasts rank 0's parameters at construction (NS/pipelines/base_pipeline.py:281-283): replicas must
            # not depend on every caller seeding identically.  The draws made on the device afterwards (pixel sampling, stratified jitter)
            # get a rank-dependent stream instead, as the reference seeds each process with seed + rank (NS/scripts/train.py:97) — with
            # identical streams every rank would train on the same pixels.
            dist.broadcast(self.flat, 0)
            if self.device.type == "cuda":
                with torch.cuda.device(self.device):
                    torch.cuda.manual_seed(torch.initial_seed() + 7919 * dist.get_rank())
        self._attach_main_grads()
        # ---- static inputs ---------------------------------------------------------------------------------------
        B, dev = self.B, self.device
        # every fp32 input is a view of ONE staging buffer, so a batch packed on the host (pack_host_batch) arrives with a single copy
        self._float_layout = [("origins", 3), ("directions", 3), ("directions_norm", 1), ("pixel_area", 1), ("rgb", 3), ("depth", 1), ("normal", 3),
                              ("jitter0", 1), ("jitter1", 1), ("jitter2", 1), ("u", 3)]
        self._float_offsets, off = {}, 0
        for name, c in self._float_layout:  # every block starts on a 16-byte boundary (vector loads), whatever B is
            self._float_offsets[name] = off
            off = (off + B * c + 3) // 4 * 4
        self._staging = torch.zeros(off, dtype=torch.float32, device=dev)
        self.inputs = {name: self._staging[self._float_offsets[name]:self._float_offsets[name] + B * c].view(B, c) for name, c in self._float_layout}
        self.inputs["camera_indices"] = torch.zeros((B, 1), dtype=torch.int64, device=dev)
        self.loss = torch.zeros((), dtype=torch.float32, device=dev)
        # proposal-weight annealing (NS/models/nerfacto.py:256-278): a device scalar the resampling kernel reads, set before every step
        self._anneal_dev = model.proposal_sampler.use_device_anneal(dev) if dev.type == "cuda" else None
        self._terms: Dict[str, torch.Tensor] = {}
        self._term_weights: Dict[str, float] = {}
        if self.device.type == "cuda" and not ops.leaf_streams.enabled:
            with torch.cuda.device(self.device):
                ops.leaf_streams.enable(3)
        self._graph_fb: Optional[torch.cuda.CUDAGraph] = None
        self._graph_opt: Optional[torch.cuda.CUDAGraph] = None
        self.launches_per_step = 0

    # MLP groups whose parameters must be contiguous in the flat buffer get ONE main-grad view on their first tensor
    def _attach_main_grads(self) -> None:
        m = self.model
        index = {id(p): v for p, v in zip(self.params, self._views)}

        def span(ps):
            a = index[id(ps[0])][0]
            last = index[id(ps[-1])]
            return self.grad[a:last[0] + last[1]]

        def attach_mlp(ps):
            # contiguity inside the flat buffer holds only when no padding was inserted between the group's tensors
            flat = ops.flat_alias([p.data for p in ps])
            if flat is None:
                repack(ps)  # odd-sized tensors: give the group its own packed storage (falls out of the single flat buffer)
                raise RuntimeError("internal: MLP group not contiguous in the flat buffer")
            ps[0]._nvo_main_grad = span(ps)

        f = m.field
        for table in [f.mlp_base.encoder.hash_table] + [pn.encoding.hash_table for pn in m.proposal_networks]:
            a, n = index[id(table)]
            table._nvo_main_grad = self.grad[a:a + n].view(table.shape)
        emb = f.embedding_appearance.embedding.weight
        a, n = index[id(emb)]
        emb._nvo_main_grad = self.grad[a:a + n].view(emb.shape)
        groups = [f.mlp_base.mlp._flat_param_list(), f.mlp_head._flat_param_list()]
        if f.use_pred_normals:
            groups.append(f.mlp_pred_normals._flat_param_list() + [f.field_head_pred_normals.net.weight, f.field_head_pred_normals.net.bias])
        for pn in m.proposal_networks:
            groups.append((pn.mlp_base[1] if not pn.use_linear else pn.linear)._flat_param_list())
        for ps in groups:
            attach_mlp(ps)

    @property
    def loss_terms(self) -> Dict[str, torch.Tensor]:
        """The weighted loss_dict entries of the last step (device scalars), built on demand."""
        return {k: v * self._term_weights[k] for k, v in self._terms.items()}

    # ---- one step --------------------------------------------------------------------------------------------------
    def _bundle(self) -> RayBundle:
        i = self.inputs
        return RayBundle(origins=i["origins"], directions=i["directions"], pixel_area=i["pixel_area"], camera_indices=i["camera_indices"],
                         metadata={"directions_norm": i["directions_norm"]})

    def _forward_backward(self, pending: bool = False) -> None:
        """zero-grad, forward, losses, backward.  From 4 ranks on (fused arm) the exchange + Adam of the "fields" group is launched from
        INSIDE the backward (ops.leaf_streams.after_field_backward), right behind the main hash-table scatter, on a high-priority
        stream; the proposal networks' backward is then ordered behind the field's chain, so both run side by side (one is NVLink
        bound, the other issue / reduction bound).  _optimizer() steps whatever has not been stepped yet."""
        i = self.inputs
        side = self.device.type == "cuda" and ops.leaf_streams.enabled
        self._fields_done = False
        self.model._before_field_forward = None
        self.model._after_field_forward = None
        ev_fill = None
        if side and self.defer_fields and pending:
            # the previous step's fields-group update next to this step's proposal sampling: optimizer (exchange) -> zero fill of the group's
            # gradient range -> this step's fp16 weight images, on the optimizer stream; the main stream joins before the field forward
            n_f = self.groups[0][2]
            if self._opt_stream is None:
                # high priority: the exchange's one CTA per SM is placed ahead of the proposal kernels pending at the same time, which fill the rest
                self._opt_stream = torch.cuda.Stream(priority=-1)
            cur = torch.cuda.current_stream()
            self._opt_stream.wait_stream(cur)
            with torch.cuda.stream(self._opt_stream):
                self._narrow_exchange = True
                try:
                    self._optimizer_group(0)
                finally:
                    self._narrow_exchange = False
                self.model.field.prepack(self.model.config.num_nerf_samples_per_ray)
                ev = torch.cuda.Event()
                ev.record(self._opt_stream)
                # the zero fill of the group's gradient range follows the exchange (peers read the gradients until its last barrier) but is
                # not in front of the field forward: the main stream joins it before the backward
                self.grad[:n_f].zero_()
                ev_fill = torch.cuda.Event()
                ev_fill.record(self._opt_stream)
            self.model._before_field_forward = lambda: torch.cuda.current_stream().wait_event(ev)
            with ops.leaf_streams.fork(self.grad):
                self.grad[n_f:].zero_()
                if self._ray_grads is not None:
                    self._ray_grads.zero_()
        elif side:
            # off the critical chain: the zero fill of the flat gradient and the fp16 weight images of the three field networks run on side
            # streams; both are joined before their first consumer.  The fill is split: the proposal networks' range (11 MB) next to the first
            # sampling kernel, the fields group's range (67 MB) behind the field forward, next to the render / loss kernels, which leave the
            # HBM idle — at the start of the step its CTAs sat in front of the first proposal forward's (launched later at the same priority)
            # and held the sampling chain back by the 18 us the fill takes (profiles/r02_timeline_final.csv vs r02_timeline_split_fill.csv)
            n_f = self.groups[0][2] if (len(self.groups) >= 2 and os.environ.get("NVO_SPLIT_FILL", "1") == "1") else 0
            with ops.leaf_streams.fork(self.grad):
                self.grad[n_f:].zero_()
                if self._ray_grads is not None:
                    self._ray_grads.zero_()
            if n_f > 0:
                def _fill_fields(n_f=n_f):
                    with ops.leaf_streams.fork(self.grad):
                        self.grad[:n_f].zero_()
                self.model._after_field_forward = _fill_fields
            self.model.field.prepack(self.model.config.num_nerf_samples_per_ray)
        else:
            self.grad.zero_()
            if self._ray_grads is not None:
                self._ray_grads.zero_()
        if side:
            for pn in self.model.proposal_networks:
                pn.preload()  # constant-memory banks of the fused proposal fields, filled next to the step's first kernels
        if self.datamanager is not None:
            if self.cam_group is not None and self.datamanager.camera_optimizer is None:
                self.datamanager.camera_optimizer = self.model.camera_optimizer  # the step prologue applies the pose correction itself
            # DynamicDataManager.next_train (nerfstudio_utils.py:295-300): pixel draws on the device unless the caller supplies them
            bundle, batch = self.datamanager.next_train(0, u=i["u"] if self.external_draws else None)
            if not self.with_normals:
                batch.pop("normal_image", None)
            if not self.external_draws:
                for k in range(3):
                    i[f"jitter{k}"].uniform_()  # the samplers' torch.rand (ray_samplers.py:115,330), on the device
        else:
            bundle = self._bundle()
            batch = {"image": i["rgb"], "depth_image": i["depth"]}
            if self.with_normals:
                batch["normal_image"] = i["normal"]
        self.model._leaf_renders = side  # proposal depth maps rendered on side streams (joined below), only inside this step
        if self._ray_grads is not None:
            ops.ray_grad_sink = (self._ray_grads[0], self._ray_grads[1])
        try:
            # eager_grads: this step differentiates `total` with total.backward() below, i.e. with grad_output == 1
            _, total, terms, weights = self.model.get_train_loss_fused(bundle, batch, [i["jitter0"], i["jitter1"], i["jitter2"]], eager_grads=True)
        finally:
            self.model._leaf_renders = False
            self.model._before_field_forward = None
            self.model._after_field_forward = None
        if side:
            ops.leaf_streams.join()  # the zero fill must have landed before the first backward kernel accumulates into the gradient
            if ev_fill is not None:
                torch.cuda.current_stream().wait_event(ev_fill)
        # Early launch of the fields group's exchange next to the (deferred) proposal backward.  Measured (profiles/r01_timeline_*s9*.csv,
        # DESIGN.md section 6): on one GPU and at 2 ranks it does not pay — both sides want the same registers and the field chain loses the
        # proposal backward that used to fill its idle issue slots; from 4 ranks on the exchange is NVLink-bound, one CTA per SM carries it, and
        # hiding it behind the proposal backward wins (8 GPUs: 1067 -> 1030 us per step).  NVO_EARLY_FIELDS_OPT=1 / 0 forces it on / off.
        env = os.environ.get("NVO_EARLY_FIELDS_OPT", "")
        multi = self.peer is not None and self.world_size >= 4
        early = (side and len(self.groups) == 2 and self.exchange != "nccl" and (env in ("1", "2") or (env != "0" and (multi or self.peer is None)))
                 and not self.defer_fields)
        # one GPU (and NVO_EARLY_FIELDS_OPT=2): the fields group's Adam (HBM-bound) starts right behind the main table scatter, which runs on
        # the high-priority critical stream, while the proposal networks' backward and scatters (issue / reduction bound) are still running
        # and stay where they are; from 4 ranks on the exchange takes that place and the proposal backward is deferred behind the field chain
        self._defer_proposal = env == "1" or (env != "2" and multi)
        if early:
            ops.leaf_streams.after_field_backward = self._fields_optimizer_hook
        self._in_backward = True
        try:
            total.backward()
        finally:
            self._in_backward = False
            ops.leaf_streams.after_field_backward = None
            ops.leaf_streams.defer_event = None
            ops.ray_grad_sink = None
        ops.clear_prepacked()
        ops.leaf_streams.join()  # scatter kernels running on side streams must land before the all-reduce / optimizer
        if self._fields_done:
            torch.cuda.current_stream().wait_stream(self._opt_stream)
        if self.defer_fields:
            self._fields_done = True  # _optimizer() leaves the fields group alone: its update opens the next step (or flush())
        self.loss.copy_(total.detach())
        if self.cam_group is not None:
            # CameraOptimizer.apply_to_raybundle's backward on the summed ray gradients (every level's sample positions + the field's direction
            # encoding), and the pose regulariser (value into the loss, gradient into the group's range of the flat gradient)
            co = self.model.camera_optimizer
            _, off, n = self.cam_group
            d_pose = self.grad[off:off + co.pose_adjustment.numel()].view(co.pose_adjustment.shape)
            if self.datamanager is not None:
                cam_idx, raw = self.datamanager._last_camera_indices, self.datamanager._last_directions_raw
            else:
                cam_idx, raw = i["camera_indices"], i["directions"]
            ops.pose_correction_backward(cam_idx.reshape(-1), raw, self._ray_grads[0], self._ray_grads[1], co.pose_adjustment.detach(), co.mode_id, d_pose=d_pose)
            ops.pose_regularizer(co.pose_adjustment.detach(), co.config.trans_l2_penalty, co.config.rot_l2_penalty, loss=self.loss, d_pose=d_pose)
        self._terms, self._term_weights = terms, weights

    def _fields_optimizer_hook(self) -> None:
        """Called by the main grid's backward on the side stream that carries the table scatter, after the scatter was launched: every
        gradient of the "fields" group is complete once that stream reaches this point."""
        cur = torch.cuda.current_stream()
        ev = torch.cuda.Event()
        ev.record(cur)
        if self._defer_proposal:
            ops.leaf_streams.defer_event = ev  # the proposal networks' backward starts here, not next to the field's backward chain
        if self._opt_stream is None:
            self._opt_stream = torch.cuda.Stream(priority=-1)
        self._opt_stream.wait_event(ev)
        with torch.cuda.stream(self._opt_stream):
            self._optimizer_group(0)
        self._fields_done = True

    def _optimizer_group(self, gi: int) -> None:
        _, off, n = self.groups[gi]
        if self.peer is not None:
            # inside the backward (early launch) the exchange shares the SMs with the proposal backward: one CTA per SM
            self.peer.adam_exchange_group(self._peer_groups[gi], self.step_counts[gi], self.lr, self.betas[0], self.betas[1], self.eps,
                                          ctas_per_sm=(int(os.environ.get("NVO_DEFER_CTAS", "1")) if getattr(self, "_narrow_exchange", False) else
                                                       1 if self._in_backward else 0))
            return
        ops.adam_step(self.flat[off:off + n], self.grad[off:off + n], self.exp_avg[off:off + n], self.exp_avg_sq[off:off + n], self.step_counts[gi],
                      self.lr, self.betas[0], self.betas[1], self.eps, 1.0 / self.world_size)

    def _optimizer(self, updated: bool = True) -> None:
        """Steps every group that has not been stepped inside the backward; the proposal group only when it received gradients."""
        if self.peer is not None and len(self.groups) == 2 and updated and not self._fields_done:
            # both groups pending at the same point of the step: one exchange launch, one pair of barriers
            self.peer.adam_exchange_groups2(self._peer_groups[0], self.step_counts[0], self._peer_groups[1], self.step_counts[1], self.lr,
                                            self.betas[0], self.betas[1], self.eps)
            self._optimizer_camera()
            return
        side = self.device.type == "cuda" and ops.leaf_streams.enabled and self.peer is None
        forked = False
        for gi, (name, _, _) in reversed(list(enumerate(self.groups))):  # the small group first: it slips in next to the big one's CTAs
            if gi == 0 and self._fields_done:
                continue
            if name == "proposal_networks" and not updated:
                continue
            if side and gi > 0:
                with ops.leaf_streams.fork():  # disjoint ranges of the flat buffers: the small group's Adam runs next to the big one's
                    self._optimizer_group(gi)
                forked = True
            else:
                self._optimizer_group(gi)
        if forked:
            ops.leaf_streams.join()
        self._fields_done = False
        self._optimizer_camera()

    def _optimizer_camera(self) -> None:
        """The "camera_opt" group: Adam under ExponentialDecayScheduler, every step (the pose deltas receive gradients through all three levels)."""
        if self.cam_group is None:
            return
        _, off, n = self.cam_group
        if self.peer is not None:
            self.peer.adam_exchange_group_decay(self._peer_cam, self.cam_step, self.cam_lr, self.cam_lr_final, self.max_num_iterations, self.betas[0],
                                                self.betas[1], self.eps, ctas_per_sm=1)
            return
        ops.adam_step_decay(self.flat[off:off + n], self.grad[off:off + n], self.exp_avg[off:off + n], self.exp_avg_sq[off:off + n], self.cam_step,
                            self.cam_lr, self.cam_lr_final, self.max_num_iterations, self.betas[0], self.betas[1], self.eps, 1.0 / self.world_size)

    def _exchange(self) -> None:
        """NCCL arm: sum-all-reduce of the flat gradient (the fused arm exchanges inside the optimizer kernel)."""
        if self.exchange == "nccl":
            sharding.allreduce_gradient_(self.grad)

    def set_inputs(self, rays: Dict[str, torch.Tensor], targets: Dict[str, torch.Tensor], jitters: Optional[List[torch.Tensor]] = None,
                   non_blocking: bool = True) -> int:
        """Copy one batch (host or device tensors) into the static input buffers; returns the bytes copied."""
        n = 0
        src = dict(rays)
        src.update({"rgb": targets["rgb"], "depth": targets["depth"], "normal": targets["normal"]})
        if jitters is None:
            for k in range(3):
                self.inputs[f"jitter{k}"].uniform_()
        else:
            for k in range(3):
                src[f"jitter{k}"] = jitters[k]
        for k, v in src.items():
            if k in self.inputs:
                self.inputs[k].copy_(v, non_blocking=non_blocking)
                n += v.numel() * v.element_size()
        return n

    def _set_sampler_state(self, updated: bool) -> None:
        """Pins ProposalNetworkSampler's `updated` predicate (ray_samplers.py:596: steps_since_update > sched(step) or step < 10)."""
        ps = self.model.proposal_sampler
        ps._step = 10 ** 6
        ps._steps_since_update = 10 ** 6 if updated else 0

    def _updated_now(self) -> bool:
        """The reference's predicate for THIS iteration; rank-invariant (depends on step counters only)."""
        if self.proposal_update == "always":
            return True
        ps = self.model.proposal_sampler
        seen = max(self.iteration - 1, 0)  # the sampler's _step is set by the AFTER_TRAIN_ITERATION callback of the previous iteration
        return bool(self._ssu > ps.update_sched(seen) or seen < 10)

    def pack_host_batch(self, rays: Dict[str, torch.Tensor], targets: Dict[str, torch.Tensor], jitters: List[torch.Tensor]):
        """One batch as two pinned host tensors (all fp32 inputs back to back in the staging layout, camera indices) for set_inputs_packed."""
        src = dict(rays)
        src.update({"rgb": targets["rgb"], "depth": targets["depth"], "normal": targets["normal"]})
        for k in range(3):
            src[f"jitter{k}"] = jitters[k]
        flat = torch.zeros(self._staging.numel(), dtype=torch.float32).pin_memory()
        for name, c in self._float_layout:
            if name not in src:
                continue  # "u": pixel draws of the dataset-fed step (pack_host_draws)
            off = self._float_offsets[name]
            flat[off:off + self.B * c].copy_(src[name].reshape(-1).float())
        cam = src["camera_indices"].reshape(self.B, 1).to(torch.int64).contiguous().pin_memory()
        return flat, cam

    def pack_host_draws(self, u: torch.Tensor, jitters: List[torch.Tensor]) -> torch.Tensor:
        """The random numbers of one dataset-fed step (pixel draws u [B,3], three jitters [B,1]) as ONE pinned host tensor in the staging
        layout, for set_draws_packed()."""
        base = self._float_offsets["jitter0"]
        flat = torch.zeros(self._staging.numel() - base, dtype=torch.float32).pin_memory()
        src = {"u": u, "jitter0": jitters[0], "jitter1": jitters[1], "jitter2": jitters[2]}
        for name, c in self._float_layout:
            if name in src:
                off = self._float_offsets[name] - base
                flat[off:off + self.B * c].copy_(src[name].reshape(-1).float())
        return flat

    def set_draws_packed(self, packed: torch.Tensor, non_blocking: bool = True) -> int:
        """One host-to-device copy: the draws of the next step (external_draws=True).  Returns the bytes copied."""
        self._staging[self._float_offsets["jitter0"]:].copy_(packed, non_blocking=non_blocking)
        return packed.numel() * 4

    def set_inputs_packed(self, packed, non_blocking: bool = True) -> int:
        """Two host-to-device copies for the whole batch; returns the bytes copied."""
        flat, cam = packed
        self._staging.copy_(flat, non_blocking=non_blocking)
        self.inputs["camera_indices"].copy_(cam, non_blocking=non_blocking)
        return flat.numel() * 4 + cam.numel() * 8

    def capture(self, warmup: int = 3) -> None:
        """Warm up on a side stream, then capture forward+backward (and the optimizer) into CUDA graphs: one graph for steps that update
        the proposal networks and, with proposal_update="reference", one for steps that do not."""
        from . import _lib

        # the capture stream gets HIGH priority, the gradient-leaf side streams (proposal backward, table scatter) keep the default (lowest):
        # kernel nodes inherit it, so when a main-chain kernel and a leaf kernel both have CTAs pending the block scheduler places the
        # main chain first and the leaf kernels fill what is left instead of pushing the chain's start back
        prio = -1 if os.environ.get("NVO_MAIN_PRIORITY", "1") == "1" else 0
        s = torch.cuda.Stream(priority=prio)
        s.wait_stream(torch.cuda.current_stream())
        modes = [True] if self.proposal_update == "always" else [True, False]
        # The warm-up (and the capture itself, which executes nothing) runs REAL steps on whatever sits in the input buffers: parameters,
        # Adam moments and the device step counters are snapshotted here and restored afterwards, so capture() is free of side effects
        # on the training state (it may be called on a fresh model or right after load_checkpoint()).
        state = self._snapshot_state()
        pendings = [False, True] if self.defer_fields else [False]
        with torch.cuda.stream(s):
            for _ in range(warmup):
                for upd in modes:
                    for pend in pendings:
                        self._set_sampler_state(upd)
                        self._forward_backward(pend)
                        self._exchange()
                        self._optimizer(upd)
            if self.defer_fields:
                self._optimizer_group(0)  # every rank leaves the warm-up with the same number of fields-group exchanges behind it
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        self._restore_state(state)
        self._pending_fields = False
        if not self.use_cuda_graph:
            return
        self._graphs = {}
        for upd in modes:
            for pend in pendings:
                self._set_sampler_state(upd)
                n0 = _lib.launch_count()
                g_fb = torch.cuda.CUDAGraph()
                g_opt = None
                with torch.cuda.graph(g_fb, stream=s):
                    self._forward_backward(pend)
                    if self.exchange != "nccl":
                        self._optimizer(upd)
                if self.exchange == "nccl":
                    g_opt = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g_opt, stream=s):
                        self._optimizer(upd)
                self._graphs[(upd, pend)] = (g_fb, g_opt)
                if upd and pend == pendings[-1]:
                    self.launches_per_step = _lib.launch_count() - n0
        self._graph_fb, self._graph_opt = self._graphs[(True, False)]

    def _moment_tensors(self) -> List[torch.Tensor]:
        if self.peer is not None:
            gids = self._peer_groups + ([self._peer_cam] if self._peer_cam is not None else [])
            return [t for gid in gids for t in self.peer.group_moments(gid)]
        return [self.exp_avg, self.exp_avg_sq]

    def full_moments(self):
        """(exp_avg, exp_avg_sq) over the whole flat buffer.  Local / NCCL arms: the buffers themselves.  Fused arm: every rank holds only
        its slice of each group, so this is a COLLECTIVE (every rank must call it) that all-gathers the slices."""
        self.flush()
        if self.peer is None:
            return self.exp_avg, self.exp_avg_sq
        from .peer import slice_range

        W = self.world_size
        outs = [torch.zeros(self.flat.numel(), dtype=torch.float32, device=self.device) for _ in range(2)]
        gids, groups = list(self._peer_groups), list(self.groups)
        if self._peer_cam is not None:
            gids.append(self._peer_cam)
            groups.append(self.cam_group)
        for gid, (_, off, n) in zip(gids, groups):
            chunk = slice_range(n, 0, W)[1]
            for j, t in enumerate(self.peer.group_moments(gid)):
                full = torch.empty(chunk * W, dtype=torch.float32, device=self.device)
                dist.all_gather_into_tensor(full, t[:chunk].contiguous())
                outs[j][off:off + n] = full[:n]
        return outs[0], outs[1]

    def load_full_moments(self, exp_avg: torch.Tensor, exp_avg_sq: torch.Tensor) -> None:
        """Inverse of full_moments(): every rank keeps (its slice of) the given whole-buffer Adam moments."""
        if self.peer is None:
            self.exp_avg.copy_(exp_avg)
            self.exp_avg_sq.copy_(exp_avg_sq)
            return
        from .peer import slice_range

        gids, groups = list(self._peer_groups), list(self.groups)
        if self._peer_cam is not None:
            gids.append(self._peer_cam)
            groups.append(self.cam_group)
        for gid, (_, off, n) in zip(gids, groups):
            lo, hi = slice_range(n, self.peer.rank, self.world_size)
            for t, src in zip(self.peer.group_moments(gid), (exp_avg, exp_avg_sq)):
                t[:hi - lo].copy_(src[off + lo:off + hi])

    def _counters(self) -> List[torch.Tensor]:
        return self.step_counts + ([self.cam_step] if self.cam_step is not None else [])

    def _snapshot_state(self):
        return ([t.clone() for t in [self.flat] + self._moment_tensors() + self._counters()], self.iteration, self._ssu,
                self.model.proposal_sampler._step, self.model.proposal_sampler._steps_since_update)

    def _restore_state(self, state) -> None:
        tensors, self.iteration, self._ssu, ps_step, ps_ssu = state
        with torch.no_grad():
            for dst, src in zip([self.flat] + self._moment_tensors() + self._counters(), tensors):
                dst.copy_(src)
        self.model.proposal_sampler._step, self.model.proposal_sampler._steps_since_update = ps_step, ps_ssu
        if self.peer is not None:
            # the step counters went back: the exchange barriers' epoch flags must follow (PeerBuffers.reset_epochs; without it the first
            # `warmup x variants` steps after capture() ran unsynchronised — tests/test_exchange.py::test_two_rank_trainer_deferred_fields_update)
            self.peer.reset_epochs()
        elif self.world_size > 1:
            torch.cuda.synchronize()
            dist.barrier()

    def flush(self) -> None:
        """Applies the deferred fields-group update of the last step (defer_fields_update=True); afterwards parameters, moments and step counters
        are exactly what the undeferred trainer holds after the same steps.  Collective at N > 1 (every rank calls it).  No-op otherwise."""
        if self._pending_fields:
            self._optimizer_group(0)
            self._pending_fields = False

    def train_step(self) -> torch.Tensor:
        """Runs one step on the current contents of the static input buffers; returns the (device) loss scalar."""
        upd = self._updated_now()
        # the reference's per-iteration callbacks: set_anneal before the step (a device scalar: graph replays read it)
        self.model.before_train_iteration(self.iteration)
        pend = self.defer_fields and self._pending_fields
        if self._graph_fb is not None:
            g_fb, g_opt = self._graphs[(upd, pend)]
            g_fb.replay()
            if self.exchange == "nccl":
                sharding.allreduce_gradient_(self.grad)
                g_opt.replay()
        else:
            from . import _lib

            n0 = _lib.launch_count()
            self._set_sampler_state(upd)
            self._forward_backward(pend)
            self._exchange()
            self._optimizer(upd)
            self.launches_per_step = _lib.launch_count() - n0
        if self.defer_fields:
            self._pending_fields = True
        # ProposalNetworkSampler.step_cb + the reset in generate_ray_samples (ray_samplers.py:591-594,611-612)
        self._ssu = 1 if upd else self._ssu + 1
        self.iteration += 1
        return self.loss
