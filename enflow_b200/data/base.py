"""Batch container, neighbour list and collater of the reference, on top of the CUDA path.

Mirrors `enflow/data/base.py:9-19` (Edges), `:21-144` (Data) and `:146-174` (DataLoader/collater):
same field names, same flat layout (molecules concatenated along axis 0, ``N`` per molecule,
``r_cut`` float32 per molecule, ``box`` tiled per atom).  ``Data.edges`` runs the K0 kernel and returns
the edge list in the reference's own order; the fused flow uses the row-grouped (CSR) form directly.
"""
import torch

from .. import _lib


class Edges:
    """`enflow/data/base.py:9-19`: (row, col), per-edge box, and the coordinates they index."""

    def __init__(self, edge_index, box, coord, csr=None):
        self.box = box
        self.row, self.col = edge_index
        self.coord = coord
        self.csr = csr      # (row32, col32, rowptr, E_dev) in CSR order, when built on the GPU

    @property
    def coord_diff(self):
        d = self.coord[self.row] - self.coord[self.col]
        half = self.box * 0.5                      # period box/2 (`base.py:18`, quirk Q8)
        return d - (d / half).round() * half


class Data:
    """`enflow/data/base.py:21-144`. Tensors may be fp64 (reference convention) or fp32."""

    def __init__(self, z=None, h=None, g=None, pos=None, vel=None, N=None, r_cut=None, box=None, label=None,
                 device='cpu'):
        self.z, self.h, self.g, self.pos, self.vel = z, h, g, pos, vel
        self.N, self.r_cut, self.box, self.label, self.device = N, r_cut, box, label, device
        self._meta = None

    # ---- layout helpers -------------------------------------------------------------------------
    def meta(self):
        """(B, mol_off int32 on the data device, max_n, N_cpu). Computed once per batch."""
        if self._meta is None:
            n_cpu = self.N.detach().to('cpu', torch.int64).reshape(-1)
            off = torch.zeros(n_cpu.numel() + 1, dtype=torch.int32)
            off[1:] = torch.cumsum(n_cpu, 0).to(torch.int32)
            self._meta = (int(n_cpu.numel()), off.to(self.pos.device), int(n_cpu.max()) if n_cpu.numel() else 0,
                          n_cpu)
        return self._meta

    def get_mol(self, i):
        if self.N.ndim == 0:
            return self
        B, off, _, n_cpu = self.meta()
        s, e = int(off[i]), int(off[i + 1])
        return Data(z=self.z[i] if self.z is not None and len(self.z) == B else self.z,
                    h=self.h[s:e], g=self.g[s:e], pos=self.pos[s:e], vel=self.vel[s:e], N=self.N[i],
                    r_cut=self.r_cut[i], box=self.box[s:e],
                    label=self.label[i] if self.label is not None and len(self.label) == B else self.label,
                    device=self.device)

    @property
    def num_atoms(self):
        return int(self.h.shape[0]) if self.N.ndim else int(self.N)

    @property
    def num_mols(self):
        return 1 if self.N.ndim == 0 else len(self.N)

    def __iter__(self):
        return (self.get_mol(i) for i in range(self.num_mols))

    def _map(self, fn, device=None):
        return Data(z=self.z, h=fn(self.h), g=fn(self.g), pos=fn(self.pos), vel=fn(self.vel), N=fn(self.N),
                    r_cut=fn(self.r_cut), box=fn(self.box), label=self.label,
                    device=self.device if device is None else device)

    def clone(self):
        return self._map(lambda t: t.clone())

    def to(self, device):
        out = self._map(lambda t: t.to(device, non_blocking=True), device=device)
        if self.N is not None and not self.N.is_cuda and self.N.ndim:
            B, off, max_n, n_cpu = self.meta()           # computed on the host: no device sync later
            out._meta = (B, off.to(device, non_blocking=True), max_n, n_cpu)
        return out

    def pin_memory(self):
        return self._map(lambda t: t.pin_memory())

    def pbc(self):
        self.pos = self.pos - (self.pos / self.box).round() * self.box      # `base.py:119-120`

    # ---- neighbour list -------------------------------------------------------------------------
    def build_edges(self, capacity=None, reference_order=True):
        """Run K0 on the current positions. Returns an Edges whose (row, col) are in the reference's
        order when ``reference_order`` (int64, `base.py:141-144`), else in row-grouped order."""
        _lib.require_cuda(self.pos)
        L = _lib.lib()
        B, off, max_n, n_cpu = self.meta()
        N = int(self.pos.shape[0])
        dev = self.pos.device
        is64 = self.pos.dtype == torch.float64
        pos = self.pos.detach().contiguous()
        box = self.box.detach().to(pos.dtype).contiguous()
        rc = self.r_cut.detach().to(dev, torch.float32).reshape(-1).contiguous()
        if capacity is None:
            capacity = int((n_cpu * (n_cpu - 1)).sum()) + 1024
        while True:
            row = torch.empty(capacity, dtype=torch.int32, device=dev)
            col = torch.empty(capacity, dtype=torch.int32, device=dev)
            refp = torch.empty(capacity, dtype=torch.int32, device=dev)
            rowptr = torch.empty(N + 1, dtype=torch.int32, device=dev)
            e_dev = torch.zeros(2, dtype=torch.int32, device=dev)
            status = torch.zeros(1, dtype=torch.int32, device=dev)
            ws = torch.empty(L.enflow_edges_workspace_ints(N), dtype=torch.int32, device=dev)
            _lib.check(L.enflow_build_edges(_lib.ptr(pos), _lib.ptr(box), int(is64), _lib.ptr(rc), _lib.ptr(off), B, N,
                                            capacity, _lib.ptr(row), _lib.ptr(col), _lib.ptr(rowptr), _lib.ptr(refp),
                                            _lib.ptr(e_dev), _lib.ptr(status), _lib.ptr(ws), _lib.stream()))
            e_used, e_total = (int(v) for v in e_dev.tolist())
            if int(status.item()) & 2:
                raise IndexError('Data.edges: fewer surviving image points than atoms '
                                 '(the reference raises IndexError at enflow/data/base.py:137)')
            if e_total <= capacity:
                break
            capacity = e_total
        row, col, refp = row[:e_total], col[:e_total], refp[:e_total]
        csr = (row, col, rowptr, e_dev)
        if reference_order:
            inv = torch.empty(e_total, dtype=torch.int64, device=dev)
            inv[refp.long()] = torch.arange(e_total, device=dev)
            r64, c64 = row.long()[inv], col.long()[inv]
        else:
            r64, c64 = row.long(), col.long()
        ebox = self.box[r64] if e_total else self.box[:0]
        return Edges(torch.stack([r64, c64]), ebox, self.pos, csr=csr)

    @property
    def edges(self):
        return self.build_edges()


class DataLoader(torch.utils.data.DataLoader):
    """`enflow/data/base.py:146-174`."""

    def __init__(self, dataset, batch_size=1, shuffle=False, **kwargs):
        super().__init__(dataset, batch_size, shuffle, collate_fn=self.collater, **kwargs)

    @staticmethod
    def collater(dataset):
        return Data(z=[d.z for d in dataset], h=torch.cat([d.h for d in dataset]),
                    g=torch.cat([d.g for d in dataset]), pos=torch.cat([d.pos for d in dataset]),
                    vel=torch.cat([d.vel for d in dataset]), N=torch.tensor([int(d.N) for d in dataset]),
                    r_cut=torch.tensor([float(d.r_cut) for d in dataset], dtype=torch.float32),
                    box=torch.cat([d.box for d in dataset]), label=[d.label for d in dataset])


def batch_from_arrays(arrs, device='cpu', dtype=torch.float64):
    """Collated Data from the dict produced by ``enflow_b200.data.synthetic.make_batch``."""
    t = lambda k: torch.as_tensor(arrs[k]).to(dtype)
    d = Data(z=None, h=t('h'), g=t('g'), pos=t('pos'), vel=t('vel'), N=torch.as_tensor(arrs['N']),
             r_cut=torch.as_tensor(arrs['r_cut'], dtype=torch.float32), box=t('box'), label=None)
    return d if device == 'cpu' else d.to(device)
