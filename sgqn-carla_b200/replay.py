"""Replay buffer with the reference's surface (utils.py:94-198: `ReplayBuffer(obs_shape, action_shape, capacity,
batch_size, prefill)`, `add`, `sample`, `sample_drq`, ...) and `LazyFrames` (utils.py:201-240), re-designed for
the GPU: single uint8 frames live once in a device-resident (or pinned-host, `storage='pinned'`) ring, a
transition is six frame-slot indices, and sampling is one fused gather(+crop/shift) kernel -- instead of the
reference's per-sample python loop, np.concatenate and pageable H2D copy on every update.
"""
import numpy as np
import torch

from ._lib import K


def _ptr(t, off=0):
    return t.data_ptr() + off * t.element_size()


class LazyFrames(object):
    """What the reference's FrameStack hands to the agent and the buffer (utils.py:201-240): k single frames that share
    storage between consecutive observations.  Here it is only a holder of the frame list -- the buffer keys its
    de-duplication on the identity of the frame arrays (`_frames`), everything else asks numpy for the stacked view."""
    __slots__ = ("_frames",)

    def __init__(self, frames, extremely_lazy=True):
        self._frames = list(frames)

    frames = property(lambda self: self._frames)
    shape = property(lambda self: (sum(f.shape[0] for f in self._frames),) + tuple(self._frames[0].shape[1:]))

    def __array__(self, dtype=None, copy=None):
        out = np.concatenate(self._frames, axis=0)
        return out if dtype is None else out.astype(dtype, copy=False)

    def __len__(self):
        return len(self._frames)

    def __getitem__(self, i):
        return np.asarray(self)[i]

    def count(self):
        return len(self._frames)

    def frame(self, i):
        return self._frames[i]


class ReplayBuffer(object):
    def __init__(self, obs_shape, action_shape, capacity, batch_size, prefill=True, device="cuda", storage="device",
                 frame_capacity=None):
        # `prefill` (utils.py:85-91 reserves host RAM) has no meaning here: the ring is allocated up front.
        assert len(obs_shape) == 3 and obs_shape[0] % 3 == 0 and obs_shape[1] == obs_shape[2]
        self.capacity, self.batch_size = int(capacity), int(batch_size)
        self.obs_shape, self.action_shape = tuple(obs_shape), tuple(action_shape)
        self.k = obs_shape[0] // 3
        assert self.k == 3, "frame_stack 3 (arguments.py:12) is what the 9-channel encoder expects"
        self.Hs = int(obs_shape[1])
        self.dev = torch.device(device)
        if storage not in ("device", "pinned"):
            raise ValueError(storage)
        self.storage = storage
        # A FrameStack rollout adds ~1 new frame per transition (consecutive stacks share 2 of 3 frames; both LazyFrames and
        # plain ndarray stacks are de-duplicated, `_slots_of`), unrelated stacks up to 6: the ring starts at 2/transition
        # and GROWS when every slot is still referenced by a live transition (`version` tells the agents that the
        # pointers baked into their CUDA graphs are stale).
        self.F = int(frame_capacity) if frame_capacity else 2 * self.capacity + 8
        self.frames = self._alloc_frames(self.F)
        self.version = 0
        self.fidx = torch.zeros(self.capacity, 6, dtype=torch.int32, device=self.dev)
        A = int(np.prod(action_shape))
        self.actions = torch.zeros(self.capacity, A, device=self.dev)
        self.rewards = torch.zeros(self.capacity, 1, device=self.dev)
        self.not_dones = torch.zeros(self.capacity, 1, device=self.dev)
        self.n_valid = torch.zeros(1, dtype=torch.int32, device=self.dev)
        self.idx, self.full = 0, False
        self._count = 0                               # transitions ever added
        self._next_frame = 0
        self._frame_last_user = np.full(self.F, -10 ** 18, dtype=np.int64)
        self._recent = {}                             # id(ndarray) -> (slot, ndarray ref) of recently uploaded frames
        self._last = []                               # [(slots, stack ndarray)] of the previous transition's obs / next_obs
        # pinned staging: frames (8 slots) and one row of transition metadata per add (64 slots); every host -> device copy
        # of add() is asynchronous, the stream is only synchronised when a staging ring wraps
        self._stage = torch.zeros(8, 3, self.Hs, self.Hs, dtype=torch.uint8).pin_memory()
        self._stage_n = 0
        self._mstage_i = torch.zeros(64, 6, dtype=torch.int32).pin_memory()
        self._mstage_f = torch.zeros(64, A + 2).pin_memory()
        self._mstage_n = 0

    def _alloc_frames(self, n):
        shape = (n, 3, self.Hs, self.Hs)
        if self.storage == "device":
            return torch.zeros(shape, dtype=torch.uint8, device=self.dev)
        return torch.zeros(shape, dtype=torch.uint8).pin_memory()   # zero-copy: the gather kernel reads host memory over PCIe / C2C

    def _grow(self):
        """Every slot belongs to a live transition: enlarge the ring by half (slot numbers stay valid)."""
        torch.cuda.synchronize(self.dev)              # nothing in flight may still read the old ring (prefetch, graphs)
        extra = max(self.F // 2, 1024)
        new = self._alloc_frames(self.F + extra)
        new[:self.F].copy_(self.frames)
        torch.cuda.synchronize(self.dev)
        self.frames = new
        self._frame_last_user = np.concatenate([self._frame_last_user, np.full(extra, -10 ** 18, dtype=np.int64)])
        self._next_frame, self.F = self.F, self.F + extra
        self.version += 1

    # ---------------------------------------------------------------- add (utils.py:111-122)
    def _live(self, slot):
        # `>=`: the transition that THIS add evicts still owns its frames until the add is complete (and, with a pinned
        # ring, a prefetch issued before the add may still be reading them)
        return self._frame_last_user[slot] >= self._count - self.capacity

    def _alloc_slot(self):
        if self._live(self._next_frame):
            self._grow()
        s = self._next_frame
        self._next_frame = (s + 1) % self.F
        return s

    def _upload(self, frame):
        frame = np.ascontiguousarray(frame, dtype=np.uint8)
        s = self._alloc_slot()
        self._frame_last_user[s] = self._count          # owned by the transition being added
        if self.storage == "device":
            if self._stage_n == self._stage.shape[0]:
                torch.cuda.current_stream().synchronize()
                self._stage_n = 0
            st = self._stage[self._stage_n]
            self._stage_n += 1
            st.copy_(torch.from_numpy(frame))
            self.frames[s].copy_(st, non_blocking=True)
        else:
            self.frames[s].copy_(torch.from_numpy(frame))
        return s

    def _slots_of(self, obs):
        lazy = getattr(obs, "_frames", None)          # our LazyFrames or the reference's (utils.py:201-240), duck-typed
        if lazy is not None:
            slots = []
            for f in lazy:
                hit = self._recent.get(id(f))
                if hit is not None and hit[1] is f and self._live(hit[0]):
                    slots.append(hit[0])
                else:
                    s = self._upload(f)
                    self._recent[id(f)] = (s, f)
                    slots.append(s)
            if len(self._recent) > 64:
                for key in list(self._recent.keys())[:-16]:
                    del self._recent[key]
            return slots, None
        arr = np.asarray(obs)
        assert arr.shape == self.obs_shape, f"obs shape {arr.shape} != {self.obs_shape}"
        # plain ndarray stacks: a frame that the previous add (or this add's obs) already uploaded is found by content --
        # obs_t is next_obs_{t-1}, and next_obs_t is obs_t shifted by one frame (FrameStack, env/wrappers.py:246-269)
        slots = [None, None, None]
        for pslots, parr in self._last:
            if all(self._live(q) for q in pslots):
                if np.array_equal(parr, arr):
                    return list(pslots), arr
                if slots[0] is None and np.array_equal(parr[3:], arr[:6]):
                    slots[0], slots[1] = pslots[1], pslots[2]
        for j in range(3):
            if slots[j] is None:
                slots[j] = self._upload(arr[3 * j:3 * j + 3])
        return slots, arr

    def add(self, obs, action, reward, next_obs, done):
        so, oarr = self._slots_of(obs)
        if oarr is not None:
            self._last = [(so, oarr)] + self._last[:1]
        sn, narr = self._slots_of(next_obs)
        self._last = [(sn, narr.copy())] if narr is not None else []
        i = self.idx
        for s in so + sn:
            self._frame_last_user[s] = self._count
        if self._mstage_n == self._mstage_i.shape[0]:
            torch.cuda.current_stream().synchronize()
            self._mstage_n = 0
        j = self._mstage_n
        self._mstage_n += 1
        A = self.actions.shape[1]
        mi, mf = self._mstage_i[j], self._mstage_f[j]
        mi.copy_(torch.tensor(so + sn, dtype=torch.int32))
        mf[:A] = torch.from_numpy(np.asarray(action, dtype=np.float32).reshape(-1))
        mf[A], mf[A + 1] = float(reward), float(not done)
        self.fidx[i].copy_(mi, non_blocking=True)
        self.actions[i].copy_(mf[:A], non_blocking=True)
        self.rewards[i].copy_(mf[A:A + 1], non_blocking=True)
        self.not_dones[i].copy_(mf[A + 1:], non_blocking=True)
        self._count += 1
        self.idx = (self.idx + 1) % self.capacity
        self.full = self.full or self.idx == 0
        self.n_valid.fill_(self.capacity if self.full else self.idx)

    def load_ring(self, frames, actions, rewards, not_dones):
        """Bulk fill from a frame ring: transition i = (frames[i:i+3], frames[i+1:i+4]) (FrameStack deque semantics,
        env/wrappers.py:240-304).  Used by the benchmarks / tests to build large synthetic buffers quickly."""
        n = len(actions)
        assert n <= self.capacity and frames.shape[0] == n + 3 and n + 3 <= self.F
        self.frames[:n + 3].copy_(torch.as_tensor(frames))
        base = torch.arange(n, dtype=torch.int32).unsqueeze(1)
        self.fidx[:n] = (base + torch.tensor([[0, 1, 2, 1, 2, 3]], dtype=torch.int32)).to(self.dev)
        self.actions[:n] = torch.as_tensor(actions).to(self.dev)
        self.rewards[:n] = torch.as_tensor(rewards).to(self.dev)
        self.not_dones[:n] = torch.as_tensor(not_dones).to(self.dev)
        self._frame_last_user[:n + 3] = n
        self._next_frame = (n + 3) % self.F
        self._count = n
        self.idx = n % self.capacity
        self.full = n == self.capacity
        self.n_valid.fill_(n)

    # ---------------------------------------------------------------- sampling
    def _get_idxs(self, n=None):
        """utils.py:124-127 (numpy global RNG, like the reference)."""
        n = self.batch_size if n is None else n
        return np.random.randint(0, self.capacity if self.full else self.idx, size=n)

    def gather_into(self, idxs_dev, obs, next_obs, action, reward, not_done, offs_dev=None, mode=0, pad=4, out_size=84):
        """Fused gather (+ random_crop mode 0 / random_shift mode 1) into caller-owned device buffers."""
        B = idxs_dev.numel()
        st = torch.cuda.current_stream().cuda_stream
        K.replay_gather(_ptr(self.frames), _ptr(self.fidx), _ptr(idxs_dev), _ptr(offs_dev) if offs_dev is not None else 0,
                        _ptr(obs), _ptr(next_obs), B, self.Hs, out_size, mode, pad, st)
        A = self.actions.shape[1]
        K.take_rows(_ptr(self.actions), _ptr(idxs_dev), _ptr(action), B, A, st)
        K.take_rows(_ptr(self.rewards), _ptr(idxs_dev), _ptr(reward), B, 1, st)
        K.take_rows(_ptr(self.not_dones), _ptr(idxs_dev), _ptr(not_done), B, 1, st)

    def _sample(self, n, mode, offs, idxs, pad=4):
        idxs = self._get_idxs(n) if idxs is None else np.asarray(idxs)
        B = len(idxs)
        out = 84 if (mode == 0 and self.Hs > 84) else self.Hs
        if mode == 0 and self.Hs > 84 and offs is None:
            cm = self.Hs - 84                                   # augmentations.py:255-256: random_(0, crop_max) exclusive
            offs = np.stack([np.random.randint(0, cm, size=(B, 2)), np.random.randint(0, cm, size=(B, 2))])
        if mode == 1 and offs is None:
            offs = np.random.randint(0, 2 * pad + 1, size=(2, B, 2))
        dev = self.dev
        idxs_dev = torch.as_tensor(idxs, dtype=torch.int64).to(dev)
        offs_dev = torch.as_tensor(np.asarray(offs), dtype=torch.int32).to(dev).contiguous() if offs is not None else None
        obs = torch.empty(B, 9, out, out, device=dev); nxt = torch.empty(B, 9, out, out, device=dev)
        a = torch.empty(B, self.actions.shape[1], device=dev); r = torch.empty(B, 1, device=dev); nd = torch.empty(B, 1, device=dev)
        self.gather_into(idxs_dev, obs, nxt, a, r, nd, offs_dev, mode, pad, out)
        return obs, a, r, nxt, nd

    def sample(self, n=None, idxs=None, offs=None):
        """utils.py:185-198: random_crop to 84 when frames are 100x100, identity at 84."""
        return self._sample(n, 0, offs, idxs)

    def sample_drq(self, n=None, pad=4, idxs=None, offs=None):
        """utils.py:158-171: random_shift(pad) on obs and next_obs."""
        return self._sample(n, 1, offs, idxs, pad)

    def sample_sacai(self, n=None, pad=4, idxs=None):
        """utils.py:173-183: no augmentation."""
        idxs = self._get_idxs(n) if idxs is None else np.asarray(idxs)
        B, dev, H = len(idxs), self.dev, self.Hs
        idxs_dev = torch.as_tensor(idxs, dtype=torch.int64).to(dev)
        obs = torch.empty(B, 9, H, H, device=dev); nxt = torch.empty(B, 9, H, H, device=dev)
        a = torch.empty(B, self.actions.shape[1], device=dev); r = torch.empty(B, 1, device=dev); nd = torch.empty(B, 1, device=dev)
        self.gather_into(idxs_dev, obs, nxt, a, r, nd, None, 0, pad, H)
        return obs, a, r, nxt, nd

    def sample_soda(self, n=None, idxs=None):
        """utils.py:137-140"""
        return self.sample_sacai(n, idxs=idxs)[0]

    def sample_curl(self, n=None, idxs=None):
        """utils.py:142-156: (obs, a, r, next_obs, nd, pos) with independent crops of the same obs."""
        idxs = self._get_idxs(n) if idxs is None else np.asarray(idxs)
        obs, a, r, nxt, nd = self._sample(None, 0, None, idxs)
        pos = self._sample(None, 0, None, idxs)[0]
        return obs, a, r, nxt, nd, pos
