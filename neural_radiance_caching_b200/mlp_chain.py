"""Host side of the tcgen05 MLP chains (csrc/chain.cu): builds the per-tile PROGRAMS
(nrc_chain_program_t) for the forward pass, the data-gradient pass and the weight-gradient
GEMMs of a Dense stack, and wraps them in one torch.autograd.Function (the role jax.custom_vjp
plays on the reference side).

A stack is described the way the reference builds it (internal/surface_light_field.py:480-500,
internal/nerf.py:461-482,561-689): concatenated inputs -> hidden Dense+ReLU layers (optionally
re-concatenating the inputs after a layer: the skip connection) -> one or more linear output
layers ("heads") reading the last activation.  Parameters stay Flax dicts {'kernel': [in,out],
'bias': [out]}; bf16 operand images are (re)packed from them on the device.
"""
import ctypes as C

import torch

from . import _lib

ATOM_BYTES = 16384
MAX_OPS, MAX_PTRS, MAX_ATOMS = 24, 40, 8
PACK_MAX, WGRAD_MAX_LAYERS, WGRAD_MAX_X, WGRAD_MAX_SEGS = 80, 12, 6, 5
OP_LOAD, OP_GEMM, OP_EPI, OP_SAVE, OP_GATHER, OP_LOADIMG = 0, 1, 2, 3, 4, 5
GEMM_ACCUMULATE, EPI_RELU, EPI_OUT_ACCUMULATE, EPI_DENSITY = 1, 1, 2, 4


class nrc_chain_op_t(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "kind", "slot", "ptr", "ld", "col0", "ncols", "npad", "tmem_col", "n", "flags", "out_ptr", "mask_ptr",
        "mask_atom0", "img_atoms", "w_chunk", "n_atoms")] + [
        ("a_slot", C.c_uint8 * MAX_ATOMS), ("a_klen", C.c_uint8 * MAX_ATOMS), ("fparam", C.c_float)]


class nrc_chain_program_t(C.Structure):
    _fields_ = [("num_ops", C.c_int32), ("slots_per_ctx", C.c_int32), ("ops", nrc_chain_op_t * MAX_OPS)]


class nrc_pack_entry_t(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("ptr", "ld", "row0", "nrows", "col0", "ncols", "chunk", "n0", "k0",
                                         "transpose")]


class nrc_wgrad_layer_t(C.Structure):
    _fields_ = [
        ("n_x_atoms", C.c_int32),
        ("x_ptr", C.c_int32 * WGRAD_MAX_X), ("x_img_atoms", C.c_int32 * WGRAD_MAX_X),
        ("x_atom", C.c_int32 * WGRAD_MAX_X), ("x_rows", C.c_int32 * WGRAD_MAX_X), ("w_row0", C.c_int32 * WGRAD_MAX_X),
        ("dy_ptr", C.c_int32), ("dy_img_atoms", C.c_int32), ("dy_atom0", C.c_int32), ("n", C.c_int32),
        ("n_seg", C.c_int32),
        ("seg_col0", C.c_int32 * WGRAD_MAX_SEGS), ("seg_ncols", C.c_int32 * WGRAD_MAX_SEGS),
        ("seg_w_ptr", C.c_int32 * WGRAD_MAX_SEGS), ("seg_b_ptr", C.c_int32 * WGRAD_MAX_SEGS),
    ]


def _pad(n, m):
    return (n + m - 1) // m * m


def _atoms_of(width):
    """[(first column, columns used)] of the 64-wide atoms covering `width` columns."""
    return [(c, min(64, width - c)) for c in range(0, width, 64)]


class ImgRef:
    """`natoms` consecutive atoms, starting at atom `atom0`, of every 128-row tile of a bf16 tile image
    [tiles][img_atoms][16 KB] (the layout SAVE writes).  Passing one where a chain takes an fp32 source / head
    gradient makes the kernel bulk-copy the atoms (NRC_OP_LOADIMG) instead of reading and converting fp32 rows;
    passing one as a destination makes the epilogue write bf16 atoms (EPI -> slot -> SAVE) instead of fp32 rows."""

    def __init__(self, img, atom0, natoms, img_atoms):
        self.img, self.atom0, self.natoms, self.img_atoms = img, atom0, natoms, img_atoms


def new_image(P, atoms, device):
    return torch.empty(_num_tiles(P) * atoms * ATOM_BYTES // 2, device=device, dtype=torch.bfloat16)


class ActImage:
    """Activations a forward chain kept for the backward pass: the tile image it saved plus, per INPUT atom, where
    that atom lives (None = in `img` at its own index; (tensor, img_atoms, atom) = in the producer's image)."""

    def __init__(self, img, in_refs):
        self.img, self.in_refs = img, in_refs

    @property
    def device(self):
        return self.img.device


class DyImage:
    """Pre-activation gradients a data-gradient chain saved, with per-head-atom overrides like ActImage."""

    def __init__(self, img, head_refs):
        self.img, self.head_refs = img, head_refs


class _Ptrs:
    """Pointer table of one launch: device tensors -> indices."""

    def __init__(self):
        self.tensors, self.index = [], {}

    def add(self, t):
        if t is None:
            return -1
        key = (t.data_ptr(), t.dtype)
        if key not in self.index:
            if len(self.tensors) >= MAX_PTRS:
                raise _lib.NrcError("chain program refers to too many buffers")
            self.index[key] = len(self.tensors)
            self.tensors.append(t)
        return self.index[key]

    def array(self):
        arr = (C.c_void_p * MAX_PTRS)()
        for i, t in enumerate(self.tensors):
            if not t.is_cuda:
                raise _lib.NrcError("nrc_b200 kernels need CUDA tensors: there is no CPU fallback")
            arr[i] = t.data_ptr()
        return arr, len(self.tensors)


class ChainSpec:
    """Static description of a Dense stack.

    in_widths : widths of the fp32 sources concatenated into the stack input (all but the last
                must be multiples of 8).
    hidden    : [(param_name, width, skip_after[, 'relu' | 'linear'])] Dense layers (ReLU unless 'linear';
                width a multiple of 16 up to 256, forward-only above 128); skip_after re-concatenates the
                stack input behind the activation (reference: `x = cat([x, inputs])`).
    heads     : [[(param_name, width), ...], ...] groups of linear output layers on the last activation;
                each group is one GEMM (sum of widths <= 128).
    """

    def __init__(self, in_widths, hidden, heads):
        self.in_widths = list(in_widths)
        for w in self.in_widths[:-1]:
            if w % 8:
                raise ValueError("all but the last concatenated input must have a width that is a multiple of 8")
        self.in_dim = sum(self.in_widths)
        self.in_pad = _pad(self.in_dim, 16)
        self.hidden_act = [(h[3] if len(h) > 3 else "relu") for h in hidden]
        for a in self.hidden_act:
            if a not in ("relu", "linear"):
                raise ValueError("hidden activations are 'relu' or 'linear'")
        self.hidden = [tuple(h[:3]) for h in hidden]
        self.heads = [[tuple(x) for x in grp] for grp in heads]
        for _, w, _ in self.hidden:
            if w % 16 or w > 256:
                raise ValueError("hidden widths must be multiples of 16, at most 256")
        self.in_atoms = _atoms_of(self.in_pad)
        self.h_slot0 = len(self.in_atoms)
        max_h = max([len(_atoms_of(w)) for _, w, _ in self.hidden] + [0])
        self.fwd_slots = self.h_slot0 + max_h
        # layer inputs: list of parts ('in' | 'h', width, padded width)
        self.x_parts = []
        parts = [("in", self.in_dim, self.in_pad)]
        for _, w, skip in self.hidden:
            self.x_parts.append(parts)
            parts = [("h", w, w)] + ([("in", self.in_dim, self.in_pad)] if skip else [])
        self.x_last = parts
        self.head_widths = [sum(w for _, w in grp) for grp in self.heads]
        for hw in self.head_widths:
            if hw > 128:
                raise ValueError("a head group is one GEMM: at most 128 output columns")
        self.head_pads = [_pad(hw, 16) for hw in self.head_widths]
        # image layouts (atoms per tile)
        self.act_atoms = len(self.in_atoms) + sum(len(_atoms_of(w)) for _, w, _ in self.hidden)
        self.dy_atoms = sum(len(_atoms_of(w)) for _, w, _ in self.hidden) + sum(
            len(_atoms_of(p)) for p in self.head_pads)
        if self.fwd_slots > 7:
            raise ValueError("stack does not fit the shared-memory slots of one tile context")
        # the data-gradient program keeps two activation-gradient tiles per context and the weight-gradient
        # kernel handles <= 128 output columns per layer: 256-wide stacks are forward-only for now
        self.supports_backward = self._bwd_slots() <= 7 and all(w <= 128 for _, w, _ in self.hidden)

    # ---- derived layouts -------------------------------------------------------------------
    def act_atom0(self, layer):
        """First atom of hidden activation `layer` inside the activation image (inputs first)."""
        a = len(self.in_atoms)
        for _, w, _ in self.hidden[:layer]:
            a += len(_atoms_of(w))
        return a

    def dy_atom0_hidden(self, layer):
        a = 0
        for _, w, _ in self.hidden[:layer]:
            a += len(_atoms_of(w))
        return a

    def dy_atom0_head(self, g):
        a = sum(len(_atoms_of(w)) for _, w, _ in self.hidden)
        for p in self.head_pads[:g]:
            a += len(_atoms_of(p))
        return a

    def _bwd_slots(self):
        head_atoms = sum(len(_atoms_of(p)) for p in self.head_pads)
        max_h = max([len(_atoms_of(w)) for _, w, _ in self.hidden] + [0])
        return max(head_atoms, max_h) + max_h   # head / current dY slots + next dY slots

    def x_atoms(self, parts, fwd_slots=True):
        """Atoms of a layer input: [(part kind, column inside the part, k extent padded to 16,
        valid columns, kernel row of the first column, forward slot)]."""
        out, row = [], 0
        for kind, w, wp in parts:
            for c, n in _atoms_of(wp):
                valid = max(0, min(n, w - c))
                slot = (self.h_slot0 if kind == "h" else 0) + c // 64
                out.append((kind, c, _pad(n, 16), valid, row + c, slot))
            row += w
        return out


class _Built:
    pass


def _build(spec):
    """Programs and tables that depend only on the spec (cached per spec)."""
    b = _Built()
    # ------------------------------------------------------------------ weight chunks
    # forward chunks: per layer, one chunk per K atom of the layer input; chunk[n][k] = W[row+k][n]
    b.pack = []          # (param_name, ld, row0, nrows, col0, ncols, chunk, n0, k0, transpose)
    chunk = 0
    b.fwd_chunk = []     # per hidden layer: [(first column, columns, first chunk)] per <=128-column block
    for li, (name, w, _) in enumerate(spec.hidden):
        blocks = []
        for nb in range(0, w, 128):
            nc = min(128, w - nb)
            blocks.append((nb, nc, chunk))
            for (_, _, _, valid, row, _) in spec.x_atoms(spec.x_parts[li]):
                if valid > 0:
                    b.pack.append((name, w, row, valid, nb, nc, chunk, 0, 0, 0))
                chunk += 1
        b.fwd_chunk.append(blocks)
    b.fwd_head_chunk = []
    for g, grp in enumerate(spec.heads):
        b.fwd_head_chunk.append(chunk)
        for (_, _, _, valid, row, _) in spec.x_atoms(spec.x_last):
            col = 0
            for name, w in grp:
                if valid > 0:
                    b.pack.append((name, w, row, valid, 0, w, chunk, col, 0, 0))
                col += w
            chunk += 1
    # data-gradient chunks: dX[part rows] = dY W^T -> B chunk[n][k] = W[row0+n][col0+k], K atoms over dY columns.
    # One GEMM op per (part, <=128-row block); its chunks are contiguous, one per dY K atom.
    def bwd_ops(parts, dy_cols_layout):
        """dy_cols_layout: [(atom first col, k extent, [(param, ld, col0 in kernel, dest k0, ncols)])]"""
        ops = []
        nonlocal chunk
        row = 0
        for kind, w, wp in parts:
            for n0 in range(0, wp, 128):
                nrows_pad = min(128, wp - n0)
                nvalid = max(0, min(nrows_pad, w - n0))
                first = chunk
                for (_, _, pieces) in dy_cols_layout:
                    for (name, ld, kcol0, k0, ncols) in pieces:
                        if nvalid > 0:
                            b.pack.append((name, ld, row + n0, nvalid, kcol0, ncols, chunk, 0, k0, 1))
                    chunk += 1
                ops.append((kind, n0, nrows_pad, first))
            row += w
        return ops

    b.bwd_hidden, b.bwd_heads = [], []
    if spec.supports_backward:
        for li, (name, w, _) in enumerate(spec.hidden):
            layout = [(c, _pad(n, 16), [(name, w, c, 0, n)]) for c, n in _atoms_of(w)]
            b.bwd_hidden.append(bwd_ops(spec.x_parts[li], layout))
        # heads: dY = concatenation of all head groups' padded columns
        head_layout = []
        for g, grp in enumerate(spec.heads):
            segs, col = [], 0
            for name, w in grp:
                segs.append((name, w, col))
                col += w
            for c, n in _atoms_of(spec.head_pads[g]):
                pieces = []
                for name, w, scol in segs:   # intersection of [scol, scol+w) with [c, c+n)
                    lo, hi = max(scol, c), min(scol + w, c + n)
                    if lo < hi:
                        pieces.append((name, w, lo - scol, lo - c, hi - lo))
                head_layout.append((c, _pad(n, 16), pieces))
        b.bwd_heads = bwd_ops(spec.x_last, head_layout)
    b.num_chunks = chunk
    if len(b.pack) > PACK_MAX:
        raise ValueError("too many weight pieces for one pack launch")
    return b


def _built(spec):
    if getattr(spec, "_built_tables", None) is None:
        spec._built_tables = _build(spec)
    return spec._built_tables


def _set_atoms(op, atoms):
    op.n_atoms = len(atoms)
    for i, (slot, klen) in enumerate(atoms):
        op.a_slot[i], op.a_klen[i] = slot, klen


def _op(prog, **kw):
    if prog.num_ops >= MAX_OPS:
        raise ValueError("chain program too long")
    op = prog.ops[prog.num_ops]
    prog.num_ops += 1
    for f in ("slot", "ptr", "out_ptr", "mask_ptr"):
        setattr(op, f, -1)
    for k, v in kw.items():
        if k == "atoms":
            _set_atoms(op, v)
        else:
            setattr(op, k, v)
    return op


def _ld(t):
    """Row stride of a 2-D fp32 view whose rows are contiguous (column slices of wider buffers are fine)."""
    if t.dim() != 2 or (t.shape[1] > 1 and t.stride(1) != 1) or t.dtype != torch.float32:
        raise _lib.NrcError("chain sources must be 2-D fp32 tensors with contiguous rows")
    return t.stride(0) if t.shape[0] > 1 else max(t.stride(0), t.shape[1])


def pack_weights_many(items, packed=None):
    """bf16 operand images (forward and data-gradient orientation) of several stacks in ONE launch.
    items: [(spec, params)].  Returns (packed buffer, [per-stack view of it])."""
    bases, total = [], 0
    for spec, _ in items:
        bases.append(total)
        total += _built(spec).num_chunks
    spec0, params0 = items[0]
    first = spec0.hidden[0][0] if spec0.hidden else spec0.heads[0][0][0]
    dev = params0[first]["kernel"].device
    if packed is None:
        packed = torch.empty(total * ATOM_BYTES // 2, device=dev, dtype=torch.bfloat16)
    ptrs = _Ptrs()
    entries = (nrc_pack_entry_t * PACK_MAX)()
    n = 0

    def flush():
        nonlocal n, ptrs
        if n:
            arr, k = ptrs.array()
            _lib.call("nrc_chain_pack_weights", _lib.stream_ptr(), entries, n, arr, k, C.c_void_p(packed.data_ptr()),
                      total, 0 if flush.first else 1)
            flush.first = False
        n, ptrs = 0, _Ptrs()

    flush.first = True
    for (spec, params), base in zip(items, bases):
        for (name, ld, row0, nrows, col0, ncols, chunk, n0, k0, tr) in _built(spec).pack:
            if n == PACK_MAX or len(ptrs.tensors) >= MAX_PTRS - 1:
                flush()
            e = entries[n]
            e.ptr = ptrs.add(params[name]["kernel"])
            e.ld, e.row0, e.nrows, e.col0, e.ncols, e.chunk, e.n0, e.k0, e.transpose = (
                ld, row0, nrows, col0, ncols, base + chunk, n0, k0, tr)
            n += 1
    flush()
    return packed, [packed[b * (ATOM_BYTES // 2):] for b in bases]


def pack_weights(spec, params):
    return pack_weights_many([(spec, params)])[1][0]


def _num_tiles(rows):
    return (rows + 127) // 128


class Batch:
    """Collects chain launches over the same rows and issues them as ONE nrc_chain_run_multi call (independent stacks:
    their tiles are dealt to the CTA pairs together).  Pass `batch=` to run_forward / run_backward_data, then flush()."""

    MAX = 3

    def __init__(self):
        self.items = []   # (program, pointer array, count, packed weights, P, keep-alive)

    def add(self, prog, ptrs, packed, P):
        arr, n = ptrs.array()
        if self.items and self.items[0][4] != P:
            raise _lib.NrcError("batched chain launches must cover the same rows")
        self.items.append((prog, arr, n, packed, P, ptrs.tensors))
        if len(self.items) == self.MAX:
            self.flush()

    def flush(self):
        if not self.items:
            return
        n = len(self.items)
        progs = (C.POINTER(nrc_chain_program_t) * n)(*[C.pointer(it[0]) for it in self.items])
        parr = (C.c_void_p * n)(*[C.cast(it[1], C.c_void_p) for it in self.items])
        narr = (C.c_int32 * n)(*[it[2] for it in self.items])
        warr = (C.c_void_p * n)(*[it[3].data_ptr() for it in self.items])
        _lib.call("nrc_chain_run_multi", _lib.stream_ptr(), n, progs, parr, narr, warr, self.items[0][4])
        self.items = []


def _launch(prog, ptrs, packed, P, batch):
    if batch is not None:
        batch.add(prog, ptrs, packed, P)
        return
    arr, n = ptrs.array()
    _lib.call("nrc_chain_run", _lib.stream_ptr(), C.byref(prog), arr, n, C.c_void_p(packed.data_ptr()), P)


def run_forward(spec, params, sources, packed, save=True, head_images=None, head_fp32=None, P=None, batch=None):
    """sources: per stack input an fp32 [P, w_i] view with contiguous rows, or an ImgRef (whole atoms: the source
    must start on a 64-column boundary).  head_images: {head group: ImgRef} - the group's result is ALSO written as
    bf16 atoms; head_fp32: {head group: False} drops the fp32 copy of such a group.  Returns (per head GROUP fp32
    buffers [P, head_pad] (None if dropped), per head layer column views of them, ActImage or None)."""
    b = _built(spec)
    first = next(t for t in sources if not isinstance(t, ImgRef)) if any(not isinstance(t, ImgRef) for t in sources) else None
    dev = first.device if first is not None else sources[0].img.device
    P = first.shape[0] if first is not None else P
    head_images = head_images or {}
    head_fp32 = head_fp32 or {}
    ptrs = _Ptrs()
    prog = nrc_chain_program_t()
    extra = max([ref.natoms for ref in head_images.values()] + [0])
    prog.slots_per_ctx = spec.fwd_slots + extra
    in_refs = [None] * len(spec.in_atoms)
    col = 0
    for i, (src, w) in enumerate(zip(sources, spec.in_widths)):
        last = i == len(sources) - 1
        if isinstance(src, ImgRef):
            wpad = (spec.in_pad - col) if last else w
            if col % 64 or src.natoms != len(_atoms_of(wpad)) or (not last and (col + w) % 64):
                raise _lib.NrcError("image sources must cover whole atoms of the stack input")
            _op(prog, kind=OP_LOADIMG, slot=col // 64, ptr=ptrs.add(src.img), col0=src.atom0, npad=src.natoms,
                img_atoms=src.img_atoms)
            for a in range(src.natoms):
                in_refs[col // 64 + a] = (src.img, src.img_atoms, src.atom0 + a)
        else:
            if src.shape[1] != w:
                raise _lib.NrcError(f"chain source {i} has width {src.shape[1]}, expected {w}")
            _op(prog, kind=OP_LOAD, slot=0, ptr=ptrs.add(src), ld=_ld(src), col0=col, ncols=w,
                npad=(spec.in_pad - col) if last else w)
        col += w
    if P is None:
        raise _lib.NrcError("run_forward needs the point count P when every source is an image")
    act = new_image(P, spec.act_atoms, dev) if save else None
    if save:
        a = 0
        while a < len(in_refs):   # input atoms that came as fp32 rows are kept for the weight gradients
            if in_refs[a] is not None:
                a += 1
                continue
            a1 = a
            while a1 < len(in_refs) and in_refs[a1] is None:
                a1 += 1
            _op(prog, kind=OP_SAVE, slot=a, ptr=ptrs.add(act), col0=a, npad=a1 - a, img_atoms=spec.act_atoms)
            a = a1
    for li, (name, w, _) in enumerate(spec.hidden):
        atoms = [(slot, klen) for (_, _, klen, _, _, slot) in spec.x_atoms(spec.x_parts[li])]
        for (nb, nc, first_chunk) in b.fwd_chunk[li]:
            _op(prog, kind=OP_GEMM, n=nc, tmem_col=nb, w_chunk=first_chunk, atoms=atoms)
        _op(prog, kind=OP_EPI, slot=spec.h_slot0, ptr=ptrs.add(params[name]["bias"]), ncols=w, npad=w, tmem_col=0,
            flags=EPI_RELU if spec.hidden_act[li] == "relu" else 0)
        if save:
            _op(prog, kind=OP_SAVE, slot=spec.h_slot0, ptr=ptrs.add(act), col0=spec.act_atom0(li),
                npad=len(_atoms_of(w)), img_atoms=spec.act_atoms)
    atoms = [(slot, klen) for (_, _, klen, _, _, slot) in spec.x_atoms(spec.x_last)]
    # head groups in batches that fit the 256 accumulator columns of a context: GEMMs of a batch, then its epilogues
    bufs, outs = [], []
    g0 = 0
    while g0 < len(spec.heads):
        g1, cols = g0, 0
        while g1 < len(spec.heads) and cols + spec.head_pads[g1] <= 256:
            cols += spec.head_pads[g1]
            g1 += 1
        tcol = 0
        for g in range(g0, g1):
            _op(prog, kind=OP_GEMM, n=spec.head_pads[g], tmem_col=tcol, w_chunk=b.fwd_head_chunk[g], atoms=atoms)
            tcol += spec.head_pads[g]
        tcol = 0
        for g in range(g0, g1):
            grp = spec.heads[g]
            want32 = head_fp32.get(g, True) or g not in head_images
            buf = torch.empty((P, spec.head_pads[g]), device=dev, dtype=torch.float32) if want32 else None
            bias = torch.cat([params[name]["bias"] for name, _ in grp]) if len(grp) > 1 else params[grp[0][0]]["bias"]
            ref = head_images.get(g)
            if ref is not None and ref.natoms != len(_atoms_of(spec.head_pads[g])):
                raise _lib.NrcError("head image must hold the whole padded head group")
            # fp32 rows leave through a staging slot (coalesced stores): after the LAST batch's GEMMs every input slot is
            # dead; an earlier batch may only use a slot its successors do not read
            if g1 == len(spec.heads):
                stage = 0
            else:
                busy = {sl for (sl, _) in atoms} | set(range(spec.fwd_slots, spec.fwd_slots + extra))
                stage = next((sl for sl in range(prog.slots_per_ctx) if sl not in busy), -1)
            _op(prog, kind=OP_EPI, slot=spec.fwd_slots if ref is not None else -1, ptr=ptrs.add(bias),
                ncols=spec.head_widths[g], npad=spec.head_pads[g], tmem_col=tcol, out_ptr=ptrs.add(buf),
                ld=spec.head_pads[g], col0=0,
                n=(stage + 1) if (buf is not None and spec.head_widths[g] >= 32) else 0)   # narrow heads: direct stores are faster
            if ref is not None:
                _op(prog, kind=OP_SAVE, slot=spec.fwd_slots, ptr=ptrs.add(ref.img), col0=ref.atom0, npad=ref.natoms,
                    img_atoms=ref.img_atoms)
            tcol += spec.head_pads[g]
            bufs.append(buf)
            c = 0
            for name, w in grp:
                outs.append(buf[:, c:c + w] if buf is not None else None)
                c += w
        g0 = g1
    _launch(prog, ptrs, packed, P, batch)
    return bufs, outs, (ActImage(act, in_refs) if save else None)


def run_density_query(spec, params, enc_desc, means, packed, head_bias, warp_c, density_bias, density, feat=None,
                      grad_pred=None, enc_out=None):
    """Fused point query on the tensor-core chain kernel (nrc_chain_query): contract + hash-grid gather as the
    tile's front end, the two hidden layers and the density / predicted-normal heads as three accumulator passes.
    spec: ChainSpec([L*F], two 64-wide hidden layers, one head group [density(1) (, pred normals(3))]);
    means [P,3]; outputs are written in place (feat [P,64], grad_pred [P,3], enc_out [P,L*F] optional)."""
    b = _built(spec)
    P = means.shape[0]
    ptrs = _Ptrs()
    prog = nrc_chain_program_t()
    prog.slots_per_ctx = 2
    _op(prog, kind=OP_GATHER, slot=0, ptr=ptrs.add(means), out_ptr=ptrs.add(enc_out), ncols=spec.in_dim,
        npad=spec.in_pad)
    for li, (name, w, _) in enumerate(spec.hidden):
        last = li == len(spec.hidden) - 1
        (nb, nc, first), = b.fwd_chunk[li]
        _op(prog, kind=OP_GEMM, n=nc, tmem_col=0, w_chunk=first, atoms=[(0, spec.in_pad)] if li == 0 else [(1, 64)])
        _op(prog, kind=OP_EPI, slot=1, ptr=ptrs.add(params[name]["bias"]), ncols=w, npad=w, tmem_col=0, flags=EPI_RELU,
            out_ptr=ptrs.add(feat) if last else -1, ld=w, col0=0)
    _op(prog, kind=OP_GEMM, n=16, tmem_col=0, w_chunk=b.fwd_head_chunk[0], atoms=[(1, 64)])
    op = _op(prog, kind=OP_EPI, slot=-1, ptr=ptrs.add(head_bias), ncols=spec.head_widths[0], npad=16, tmem_col=0,
             flags=EPI_DENSITY, out_ptr=ptrs.add(density), mask_ptr=ptrs.add(grad_pred))
    op.fparam = float(density_bias)
    arr, n = ptrs.array()
    _lib.call("nrc_chain_query", _lib.stream_ptr(), C.byref(prog), arr, n, C.c_void_p(packed.data_ptr()), P,
              C.byref(enc_desc), float(warp_c))


def run_backward_data(spec, params, g_heads, act, packed, P, d_src=None, batch=None):
    """Data-gradient pass.  g_heads: per head GROUP an fp32 [P, >= group width] view (rows contiguous) or an ImgRef
    holding the group's padded columns as bf16 atoms.  d_src: None (no input gradient) or per source one of
    None / (fp32 [P, w_i] view, accumulate flag) / ImgRef (bf16 atoms; the source must start on a 64-column
    boundary).  The gradient of the stack input is accumulated over the skip connections in tensor memory
    (accumulator columns beside the working ones) and leaves the SM once.  Returns the DyImage for the weight
    gradients."""
    if not spec.supports_backward:
        raise NotImplementedError("this stack is forward-only (hidden width > 128)")
    b = _built(spec)
    act_img = act.img if isinstance(act, ActImage) else act
    dev = act_img.device
    dy = new_image(P, spec.dy_atoms, dev)
    ptrs = _Ptrs()
    prog = nrc_chain_program_t()
    max_h = max([len(_atoms_of(w)) for _, w, _ in spec.hidden] + [0])
    head_atoms_n = sum(len(_atoms_of(p)) for p in spec.head_pads)
    out_atoms = 0
    if d_src is not None:
        c = 0
        for i, w in enumerate(spec.in_widths):
            if isinstance(d_src[i], ImgRef):
                last = i == len(spec.in_widths) - 1
                wpad = (spec.in_pad - c) if last else w
                if c % 64 or d_src[i].natoms != len(_atoms_of(wpad)):
                    raise _lib.NrcError("image destinations must cover whole atoms of the stack input")
                out_atoms = max(out_atoms, d_src[i].natoms)
            c += w
    n_extra = sum(len(_atoms_of(spec.head_pads[g])) for g, gb in enumerate(g_heads)
                  if isinstance(gb, (list, tuple)) for _ in gb[1:])
    head_atoms_n += n_extra
    S = max(spec._bwd_slots() + n_extra, max(head_atoms_n, max_h) + out_atoms)
    if S > 7:
        raise ValueError("data-gradient program does not fit the shared-memory slots of one tile context")
    prog.slots_per_ctx = S
    cur0, nxt0 = 0, S - max(max_h, out_atoms)
    slot = 0
    head_atoms, head_refs = [], []
    extra_sets = []   # (head group, first slot): further images summed into the group's gradient (dX is linear in dY)
    for g, gbuf in enumerate(g_heads):
        hp = spec.head_pads[g]
        na = len(_atoms_of(hp))
        if isinstance(gbuf, (list, tuple)):
            gbuf, more = gbuf[0], list(gbuf[1:])
        else:
            more = []
        if isinstance(gbuf, ImgRef):
            if gbuf.natoms != na:
                raise _lib.NrcError("head-gradient image must hold the whole padded head group")
            _op(prog, kind=OP_LOADIMG, slot=slot, ptr=ptrs.add(gbuf.img), col0=gbuf.atom0, npad=na, img_atoms=gbuf.img_atoms)
            head_refs += [(gbuf.img, gbuf.img_atoms, gbuf.atom0 + a) for a in range(na)]
            for ref in more:
                extra_sets.append((g, ref))
        else:
            _op(prog, kind=OP_LOAD, slot=slot, ptr=ptrs.add(gbuf), ld=_ld(gbuf), col0=0, ncols=spec.head_widths[g], npad=hp)
            _op(prog, kind=OP_SAVE, slot=slot, ptr=ptrs.add(dy), col0=spec.dy_atom0_head(g), npad=na, img_atoms=spec.dy_atoms)
            head_refs += [None] * na
        for c, n in _atoms_of(hp):
            head_atoms.append((slot + c // 64, _pad(n, 16)))
        slot += na
    extra_gemms = []   # (first K atom inside the head layout, [(slot, klen)])
    for g, ref in extra_sets:
        hp = spec.head_pads[g]
        na = len(_atoms_of(hp))
        if ref.natoms != na:
            raise _lib.NrcError("head-gradient image must hold the whole padded head group")
        _op(prog, kind=OP_LOADIMG, slot=slot, ptr=ptrs.add(ref.img), col0=ref.atom0, npad=na, img_atoms=ref.img_atoms)
        k0 = sum(len(_atoms_of(p)) for p in spec.head_pads[:g])
        extra_gemms.append((k0, [(slot + c // 64, _pad(n, 16)) for c, n in _atoms_of(hp)]))
        slot += na
    if d_src is not None:
        c = 0
        for w in spec.in_widths:
            if c % 16:
                raise ValueError("input gradients need 16-aligned source offsets")
            c += w
    # accumulator columns of the input gradient: beside the working columns [0, 128)
    acc_in0 = 128 if spec.in_pad <= 128 else 256
    in_started = [False]

    def emit(parts_ops, a_atoms, mask_layer, out_slot0, extra=()):
        """GEMMs of one layer's data gradient: the input part accumulates in its own tensor-memory columns, the
        hidden part becomes the next dY (ReLU mask of the layer below applied by the epilogue).  extra: further
        operand sets (first K atom, atoms) multiplied with the same weights and accumulated."""
        if d_src is not None:
            for (_, n0, npad, first) in [o for o in parts_ops if o[0] == "in"]:
                _op(prog, kind=OP_GEMM, n=npad, tmem_col=acc_in0 + n0, w_chunk=first, atoms=a_atoms,
                    flags=GEMM_ACCUMULATE if in_started[0] else 0)
                for (k0, atoms) in extra:
                    _op(prog, kind=OP_GEMM, n=npad, tmem_col=acc_in0 + n0, w_chunk=first + k0, atoms=atoms, flags=GEMM_ACCUMULATE)
            if any(o[0] == "in" for o in parts_ops):
                in_started[0] = True
        ops = [o for o in parts_ops if o[0] == "h"]
        if not ops:
            return
        for (_, n0, npad, first) in ops:
            _op(prog, kind=OP_GEMM, n=npad, tmem_col=n0, w_chunk=first, atoms=a_atoms)
            for (k0, atoms) in extra:
                _op(prog, kind=OP_GEMM, n=npad, tmem_col=n0, w_chunk=first + k0, atoms=atoms, flags=GEMM_ACCUMULATE)
        w = spec.hidden[mask_layer][1]
        if spec.hidden_act[mask_layer] == "relu":
            _op(prog, kind=OP_EPI, slot=out_slot0, ncols=w, npad=w, tmem_col=0, mask_ptr=ptrs.add(act_img),
                mask_atom0=spec.act_atom0(mask_layer), img_atoms=spec.act_atoms)
        else:
            _op(prog, kind=OP_EPI, slot=out_slot0, ncols=w, npad=w, tmem_col=0)
        _op(prog, kind=OP_SAVE, slot=out_slot0, ptr=ptrs.add(dy), col0=spec.dy_atom0_hidden(mask_layer),
            npad=len(_atoms_of(w)), img_atoms=spec.dy_atoms)

    nh = len(spec.hidden)
    emit(b.bwd_heads, head_atoms, nh - 1, nxt0, extra_gemms)
    cur0, nxt0 = nxt0, cur0
    for li in range(nh - 1, -1, -1):
        w = spec.hidden[li][1]
        a_atoms = [(cur0 + c // 64, _pad(n, 16)) for c, n in _atoms_of(w)]
        emit(b.bwd_hidden[li], a_atoms, li - 1, nxt0)
        cur0, nxt0 = nxt0, cur0
    if d_src is not None and in_started[0]:
        out0 = S - out_atoms
        c = 0
        for i, w in enumerate(spec.in_widths):
            dst = d_src[i]
            last = i == len(spec.in_widths) - 1
            if isinstance(dst, ImgRef):
                wpad = (spec.in_pad - c) if last else w
                _op(prog, kind=OP_EPI, slot=out0, ncols=_pad(wpad, 16), npad=_pad(wpad, 16), tmem_col=acc_in0 + c)
                _op(prog, kind=OP_SAVE, slot=out0, ptr=ptrs.add(dst.img), col0=dst.atom0, npad=dst.natoms,
                    img_atoms=dst.img_atoms)
            elif dst is not None and dst[0] is not None:
                t, accumulate = dst
                _op(prog, kind=OP_EPI, slot=-1, ncols=w, npad=_pad(w, 16), tmem_col=acc_in0 + c, out_ptr=ptrs.add(t),
                    ld=_ld(t), col0=0, flags=EPI_OUT_ACCUMULATE if accumulate else 0,
                    n=1 if ((out_atoms == 0 or out0 > 0) and w >= 32) else 0)   # staged through slot 0 (every GEMM is done)
            c += w
    _launch(prog, ptrs, packed, P, batch)
    return DyImage(dy, head_refs)


def wgrad_layers(spec, act, dy, grad_sinks, wptrs, extra_head_dy=None):
    """Weight-gradient descriptors of one stack; grad_sinks[name] = (g_kernel, g_bias) accumulated into.
    act: ActImage, dy: DyImage.  extra_head_dy: {head group: [ImgRef, ...]} further dY images of a head group whose
    upstream gradient arrives as a SUM of several images (each adds one more descriptor: dW is linear in dY)."""
    act_img = act.img if isinstance(act, ActImage) else act
    in_refs = act.in_refs if isinstance(act, ActImage) else [None] * len(spec.in_atoms)
    dy_img = dy.img if isinstance(dy, DyImage) else dy
    head_refs = dy.head_refs if isinstance(dy, DyImage) else None
    ia, idy = wptrs.add(act_img), wptrs.add(dy_img)
    nh = len(spec.hidden)
    layers = []

    def fill_x(L, parts, layer_index):
        atoms = [a for a in spec.x_atoms(parts) if a[3] > 0]
        if len(atoms) > WGRAD_MAX_X:
            raise ValueError("layer input too wide for one weight-gradient pass")
        L.n_x_atoms = len(atoms)
        for i, (kind, c, _, valid, row, _) in enumerate(atoms):
            ref = in_refs[c // 64] if kind == "in" else None
            if ref is not None:
                L.x_ptr[i], L.x_img_atoms[i], L.x_atom[i] = wptrs.add(ref[0]), ref[1], ref[2]
            else:
                L.x_ptr[i], L.x_img_atoms[i] = ia, spec.act_atoms
                L.x_atom[i] = (spec.act_atom0(layer_index - 1) if kind == "h" else 0) + c // 64
            L.x_rows[i], L.w_row0[i] = valid, row

    for li, (name, w, _) in enumerate(spec.hidden):
        L = nrc_wgrad_layer_t()
        fill_x(L, spec.x_parts[li], li)
        L.dy_ptr, L.dy_img_atoms, L.dy_atom0, L.n = idy, spec.dy_atoms, spec.dy_atom0_hidden(li), w
        L.n_seg = 1
        gk, gb = grad_sinks[name]
        L.seg_col0[0], L.seg_ncols[0], L.seg_w_ptr[0], L.seg_b_ptr[0] = 0, w, wptrs.add(gk), wptrs.add(gb)
        layers.append(L)
    ha = 0
    for g, grp in enumerate(spec.heads):
        na = len(_atoms_of(spec.head_pads[g]))
        ref = head_refs[ha] if head_refs is not None else None
        srcs = [(idy, spec.dy_atoms, spec.dy_atom0_head(g))] if ref is None else [(wptrs.add(ref[0]), ref[1], ref[2])]
        for extra in (extra_head_dy or {}).get(g, []):
            srcs.append((wptrs.add(extra.img), extra.img_atoms, extra.atom0))
        ha += na
        for (p_dy, dy_atoms, dy_atom0) in srcs:
            L = nrc_wgrad_layer_t()
            fill_x(L, spec.x_last, nh)
            L.dy_ptr, L.dy_img_atoms, L.dy_atom0, L.n = p_dy, dy_atoms, dy_atom0, spec.head_pads[g]
            L.n_seg = len(grp)
            col = 0
            for s, (name, w) in enumerate(grp):
                gk, gb = grad_sinks[name]
                L.seg_col0[s], L.seg_ncols[s], L.seg_w_ptr[s], L.seg_b_ptr[s] = col, w, wptrs.add(gk), wptrs.add(gb)
                col += w
            layers.append(L)
    return layers


def wgrad_launch(layers, wptrs, P):
    warr, wn = wptrs.array()
    for i in range(0, len(layers), WGRAD_MAX_LAYERS):
        part = layers[i:i + WGRAD_MAX_LAYERS]
        carr = (nrc_wgrad_layer_t * len(part))(*part)
        _lib.call("nrc_chain_wgrad", _lib.stream_ptr(), carr, len(part), warr, wn, P)


def resolve_sinks(named_params):
    """named_params: {name: (kernel, bias)} -> ({name: (g_kernel, g_bias)}, {name: sunk?}).  Registered
    gradient sinks (_lib.register_grad_sink) are accumulated into directly; otherwise fresh zeros."""
    sinks, sunk = {}, {}
    for name, (kern, bias) in named_params.items():
        sk, sb = _lib.grad_sink(kern), _lib.grad_sink(bias)
        if sk is not None and sb is not None:
            sinks[name], sunk[name] = (sk, sb), True
        else:
            sinks[name], sunk[name] = (torch.zeros_like(kern), torch.zeros_like(bias)), False
    return sinks, sunk


class _ChainFn(torch.autograd.Function):
    """custom_vjp of a whole Dense stack."""

    @staticmethod
    def forward(ctx, spec, names, n_src, *tensors):
        sources = [t if (t.dim() == 2 and t.stride(-1) == 1) else t.contiguous() for t in tensors[:n_src]]
        flat = tensors[n_src:]
        params = {name: {"kernel": flat[2 * i], "bias": flat[2 * i + 1]} for i, name in enumerate(names)}
        need_grad = any(t.requires_grad for t in tensors)
        if need_grad and not spec.supports_backward:
            raise NotImplementedError("this stack is forward-only (hidden width > 128): call it under torch.no_grad()")
        packed = pack_weights(spec, params)
        _, outs, act = run_forward(spec, params, sources, packed, save=need_grad)
        ctx.spec, ctx.names, ctx.n_src = spec, names, n_src
        ctx.P = sources[0].shape[0]
        ctx.save_for_backward(act.img if act is not None else None, packed, *flat)
        ctx.src_needs = [t.requires_grad for t in tensors[:n_src]]
        return tuple(o.contiguous() for o in outs)

    @staticmethod
    def backward(ctx, *g_outs):
        spec, names, n_src = ctx.spec, ctx.names, ctx.n_src
        act, packed, *flat = ctx.saved_tensors
        act = ActImage(act, [None] * len(spec.in_atoms))
        params = {name: {"kernel": flat[2 * i], "bias": flat[2 * i + 1]} for i, name in enumerate(names)}
        P = ctx.P
        dev = act.device
        g_heads, k = [], 0
        for g, grp in enumerate(spec.heads):
            buf = torch.zeros((P, spec.head_pads[g]), device=dev, dtype=torch.float32)
            c = 0
            for name, w in grp:
                if g_outs[k] is not None:
                    buf[:, c:c + w] = g_outs[k]
                c += w
                k += 1
            g_heads.append(buf)
        sinks, sunk = resolve_sinks({name: (flat[2 * i], flat[2 * i + 1]) for i, name in enumerate(names)})
        d_src = None
        if any(ctx.src_needs):
            d_src = [(torch.empty((P, w), device=dev, dtype=torch.float32), False) for w in spec.in_widths]
        dy = run_backward_data(spec, params, g_heads, act, packed, P, d_src)
        wptrs = _Ptrs()
        wgrad_launch(wgrad_layers(spec, act, dy, sinks, wptrs), wptrs, P)
        grads = [None, None, None]
        for i in range(n_src):
            grads.append(d_src[i][0] if (d_src is not None and ctx.src_needs[i]) else None)
        for name in names:
            grads += [None, None] if sunk[name] else list(sinks[name])
        return tuple(grads)


_param_generation = [0]


def bump_param_generation():
    """Tell every PackCache that parameters may have changed WITHOUT torch noticing: updates through raw device pointers
    (an external optimizer, C-ABI kernels, in-place updates replayed inside a CUDA graph) do not bump `_version`.  A
    training harness calls this once per optimizer step (workload.CacheTrainStep does)."""
    _param_generation[0] += 1


class PackCache:
    """Packed bf16 operand images keyed by the parameters' identity, torch version counters and the parameter
    generation (bump_param_generation): a render loop packs once per parameter update instead of once per chunk.  The
    keyed tensors are held, so a freed parameter's address cannot be recycled by a different tensor that would then hit
    the stale entry.  invalidate() drops the entry explicitly (e.g. after loading a checkpoint into the same storage)."""

    def __init__(self):
        self._store = {}

    def invalidate(self):
        self._store.clear()

    def get(self, key_tensors, build):
        key = (_param_generation[0],) + tuple((id(t), t.data_ptr(), t._version) for t in key_tensors)
        hit = self._store.get("k")
        if hit is None or hit[0] != key:
            self._store["k"] = (key, build(), list(key_tensors))
        return self._store["k"][1]


def forward_cached(spec, params, sources, cache):
    """Forward-only evaluation (no autograd) with the packed weights taken from `cache` (PackCache)."""
    names = [h[0] for h in spec.hidden] + [name for grp in spec.heads for name, _ in grp]
    keys = [params[n]["kernel"] for n in names]
    packed = cache.get(keys, lambda: pack_weights(spec, params))
    srcs = [t if (t.dim() == 2 and t.stride(-1) == 1) else t.contiguous() for t in sources]
    _, outs, _ = run_forward(spec, params, srcs, packed, save=False)
    return outs


def apply(spec, params, sources):
    """Run the stack; returns one fp32 [P, w] tensor per output layer (in spec.heads order)."""
    names = [h[0] for h in spec.hidden] + [name for grp in spec.heads for name, _ in grp]
    flat = []
    for name in names:
        flat += [params[name]["kernel"], params[name]["bias"]]
    return _ChainFn.apply(spec, tuple(names), len(sources), *sources, *flat)
