"""One training step of the VAE as a single CUDA-graph replay.

The reference's loop body (main.py:389-390,397-398)

    mu, logvar, enc, recon = model(image);  loss, *_ = model.loss(target, mu, logvar, enc, recon, device, args)
    optimizer.zero_grad();  loss.backward()

issues ~160 short sm_100a kernels; at a few microseconds each the host cannot launch them as fast as a
B200 retires them.  `GraphedTrainStep` runs exactly that body once under stream capture (through the same
public `VAE.forward` / `VAE.loss` / autograd path, so it is the same code that the parity tests check) and
replays the captured graph afterwards: one launch per step.  Shapes are fixed by construction; inputs are
copied into static buffers, outputs and `.grad` tensors are static tensors that every replay overwrites.

The rsample noise stays fresh across replays: the Philox {seed, offset} pair lives in device memory and is
advanced inside the graph (include/mmvae.h, `rng_state`).  The KL weight is a device scalar too
(`set_kl_weight`), so a schedule can anneal it between replays, and an optional `mmvae_b200.FusedAdam` is
captured behind the backward (its step count lives on the device), which makes one replay the reference's whole
loop body main.py:389-399.

Side effects on the module are confined to what a real step has: the warm-up steps that precede the capture
run on a snapshot of the BatchNorm buffers / counters / noise offset that is restored afterwards, `defer_metrics`
is restored after the capture, and every replay re-attaches the graph's static gradient tensors to `p.grad`
(an `optimizer.zero_grad()` or a second GraphedTrainStep on the same module may have detached them).
"""
import types

import torch


class GraphedTrainStep:
    def __init__(self, model, batch_size, args=None, kl_weight=None, warmup=3, from_labels=None, optimizer=None):
        """`from_labels=(data_mean, data_std)`: the step starts from the uint8 k-means label map
        (`self.labels`, [N,S,S]) and normalises it on the device (main.py:381-388) inside the graph.
        `kl_weight`: initial KL coefficient (default: the module's `kl`); change it with `set_kl_weight`.
        `optimizer`: a mmvae_b200.FusedAdam over `model` whose step is captured behind the backward."""
        if not next(model.parameters()).is_cuda:
            raise ValueError("GraphedTrainStep needs the module on a CUDA device")
        self.model = model
        dev = next(model.parameters()).device
        s = model.input_image_size
        categorical = model.decoder_out_channels > model.in_channels
        self.x = torch.zeros(batch_size, model.in_channels, s, s, dtype=torch.float32, device=dev)
        self.target = (torch.zeros(batch_size, s, s, dtype=torch.int64, device=dev) if categorical else self.x)
        self.args = args if args is not None else types.SimpleNamespace(data_ratio_of_labels=None)
        self._one = torch.ones((), dtype=torch.float32, device=dev)      # d loss / d loss: no fill kernel inside the graph
        self.kl_weight_dev = torch.tensor(float(model.kl if kl_weight is None else kl_weight), dtype=torch.float32, device=dev)
        self.optimizer = optimizer
        self._own_target = categorical
        self.from_labels = from_labels
        self.labels = torch.zeros(batch_size, s, s, dtype=torch.uint8, device=dev) if from_labels is not None else None
        model.train(True)
        saved_defer = model.defer_metrics
        model.defer_metrics = True                      # no host synchronisation inside the step
        if model.require_rsample and model._rng_dev is None:
            if model._philox_seed is None:
                model._philox_seed = torch.initial_seed() & 0xFFFFFFFFFFFFFFFF
            seed = model._philox_seed
            seed = seed - (1 << 64) if seed >= (1 << 63) else seed
            model._rng_dev = torch.tensor([seed, model._philox_offset], dtype=torch.int64, device=dev)

        # warm-up steps (they allocate the workspace, load the kernels, create the autograd nodes) must not leave a
        # trace: snapshot everything a training-mode step mutates and put it back before the capture
        snap = [t.clone() for t in (model._arena, model._bn_arena, model._counters)]
        rng_snap = model._rng_dev.clone() if model._rng_dev is not None else None
        host_rng = (model._philox_offset, model._mmd_calls)
        opt_snap = None
        if optimizer is not None:
            opt_snap = (optimizer.exp_avg.clone(), optimizer.exp_avg_sq.clone(), optimizer._step_dev.clone(), optimizer.step_count)
        # one warm-up / capture stream per module: the parameters' AccumulateGrad nodes live on the stream that was current
        # when they were created (and a GraphedTrainStep keeps them alive through its outputs), so a second graph of the
        # same module captured on another stream would make autograd insert a cross-stream hand-over per parameter
        side = getattr(model, "_capture_stream", None)
        if side is None or side.device != dev:
            side = model._capture_stream = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):
                self._body()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)

        def restore():
            with torch.no_grad():
                for t, s in zip((model._arena, model._bn_arena, model._counters), snap):
                    t.copy_(s)
                if rng_snap is not None:
                    model._rng_dev.copy_(rng_snap)
                model._philox_offset, model._mmd_calls = host_rng
                if opt_snap is not None:
                    optimizer.exp_avg.copy_(opt_snap[0]); optimizer.exp_avg_sq.copy_(opt_snap[1])
                    optimizer._step_dev.copy_(opt_snap[2]); optimizer.step_count = opt_snap[3]

        restore()
        for p in model.parameters():
            p.grad = None
        self.graph = torch.cuda.CUDAGraph()
        # captured on the stream the warm-up ran on: the parameters' AccumulateGrad nodes were created there, and a
        # capture on another stream makes autograd warn about (and insert) a cross-stream hand-over per parameter
        with torch.cuda.graph(self.graph, stream=side):
            self._body()
        torch.cuda.synchronize(dev)
        restore()                                       # the capture itself ran host-side bookkeeping once
        self._params = list(model.parameters())
        self.grads = [p.grad for p in self._params]
        model.defer_metrics = saved_defer
        model._kl_dev = None
        self._copy_stream = None                        # prefetch(): double-buffered H2D beside the running step
        self._staged = None

    def close(self):
        """Release the captured graph (and the NCCL work a data-parallel capture holds on the communicator: a process
        group must not be destroyed while such a graph is alive)."""
        self.graph = None

    def set_kl_weight(self, value):
        """KL annealing: the coefficient of the next replays (a device scalar the captured loss kernels read)."""
        self.kl_weight_dev.fill_(float(value))

    def _body(self):
        m = self.model
        m._kl_dev = self.kl_weight_dev
        for p in m.parameters():
            p.grad = None                               # optimizer.zero_grad(), main.py:397
        if self.from_labels is not None:
            from .data import prepare_input
            prepare_input(self.labels, self.from_labels[0], self.from_labels[1], want_target=self._own_target,
                          out=self.x, out_target=self.target if self._own_target else None)
        mu, logvar, enc, recon = m(self.x)
        m._mmd_join_deferred = True                     # the MMD diagnostic's side branch joins at the end of the step
        try:
            loss, pxz, kl, mmd = m.loss(self.target, mu, logvar, enc, recon, self.x.device, self.args)
        finally:
            m._mmd_join_deferred = False
        loss.backward(self._one)
        if self.optimizer is not None:
            self.optimizer.step(m.last_flat_grad)       # optimizer.step(), main.py:399
        if m._mmd_pending:
            m._mmd_join()
            m._mmd_pending = False
        self.mu, self.logvar, self.encoding, self.reconstruction = mu, logvar, enc, recon
        self.loss, self.pxz, self.kl, self.mmd = loss.detach(), pxz, kl, mmd

    def prefetch(self, x):
        """Start the host-to-device copy of the NEXT batch (uint8 label map with `from_labels`, else the fp32 input;
        pinned host memory) on a copy stream into one of two staging buffers, so that it overlaps the step that is
        running; the next `step()` call without arguments consumes it (a 1 MB device-to-device copy in front of the replay).
        The input side of the loop (main.py:374-388) as a double-buffered pipeline."""
        buf = self.labels if self.labels is not None else self.x
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=buf.device)
            self._stage = [torch.empty_like(buf) for _ in range(2)]
            self._ready = [torch.cuda.Event() for _ in range(2)]
            self._consumed = [None, None]
            self._slot = 0
        s = self._slot
        self._slot ^= 1
        cs = self._copy_stream
        if self._consumed[s] is not None:
            cs.wait_event(self._consumed[s])            # the step that read this staging buffer has copied it out
        with torch.cuda.stream(cs):
            self._stage[s].copy_(x, non_blocking=True)
            self._ready[s].record(cs)
        self._staged = s

    def __call__(self, x=None, target=None):
        """Copy the batch into the static input (pass None when `self.x` was filled in place or a batch was staged by
        `prefetch`) and replay.
        Returns (loss, pxz/N, KL/N) as 0-d device tensors; gradients are in `p.grad` of every parameter."""
        if x is None and self._staged is not None:
            s, self._staged = self._staged, None
            cur = torch.cuda.current_stream()
            cur.wait_event(self._ready[s])
            (self.labels if self.labels is not None else self.x).copy_(self._stage[s], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(cur)
            self._consumed[s] = ev
        if x is not None and self.labels is not None:
            if x.data_ptr() != self.labels.data_ptr():
                self.labels.copy_(x, non_blocking=True)     # H2D straight into the static buffer when x is pinned host memory
        elif x is not None and x.data_ptr() != self.x.data_ptr():
            self.x.copy_(x, non_blocking=True)
        if self._own_target and target is not None:
            self.target.copy_(target, non_blocking=True)
        for p, g in zip(self._params, self.grads):      # zero_grad(set_to_none) / another graph may have detached them
            pg = p.grad
            if pg is None or pg.data_ptr() != g.data_ptr():
                p.grad = g
        self.graph.replay()
        if self.optimizer is not None:
            self.optimizer.step_count += 1              # the device-side count advanced inside the graph
        return self.loss, self.pxz, self.kl
