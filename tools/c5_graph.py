"""Synthetic recommender-shaped graph of BASELINE config 5, generated on the device with torch:
users x items, each entry Bernoulli(dens) (Poisson row degrees, uniform columns, duplicates removed), or a
Pareto(1.5) user-degree variant with the same mean; CSR of Y and of Y' with sorted rows, int32."""
import torch


def make_graph(ns, nt, dens, dev, degrees="poisson", weighted=False, seed=20245):
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    if degrees == "pareto":
        u = torch.rand(ns, device=dev, generator=g).clamp_min(1e-12)
        deg = ((nt * dens / 3.0) * u.pow(-1.0 / 1.5)).clamp_max(20000.0).to(torch.int64)
    else:
        deg = torch.poisson(torch.full((ns,), nt * dens, device=dev), generator=g).to(torch.int64)
    rows = torch.repeat_interleave(torch.arange(ns, device=dev), deg)
    cols = torch.randint(0, nt, (rows.numel(),), device=dev, generator=g)
    keys = torch.unique(rows * nt + cols)  # sorted: by row, then column
    rows, cols = keys // nt, keys % nt
    nnz = keys.numel()
    y_ptr = torch.zeros(ns + 1, dtype=torch.int32, device=dev)
    y_ptr[1:] = torch.cumsum(torch.bincount(rows, minlength=ns), 0).to(torch.int32)
    y_idx = cols.to(torch.int32)
    keyt, perm = torch.sort(cols * ns + rows)
    yt_ptr = torch.zeros(nt + 1, dtype=torch.int32, device=dev)
    yt_ptr[1:] = torch.cumsum(torch.bincount(keyt // ns, minlength=nt), 0).to(torch.int32)
    yt_idx = (keyt % ns).to(torch.int32)
    y_val = (torch.rand(nnz, dtype=torch.float64, device=dev, generator=g) + 0.5) if weighted else None
    yt_val = y_val[perm].contiguous() if weighted else None
    del keys, keyt, rows, cols, perm
    torch.cuda.synchronize()
    return dict(ns=ns, nt=nt, nnz=nnz, y_ptr=y_ptr, y_idx=y_idx, y_val=y_val, yt_ptr=yt_ptr, yt_idx=yt_idx, yt_val=yt_val)


def wrap(lib, ctx, check, G):
    """ss_csr handles over the torch buffers of make_graph (no copy)."""
    import ctypes as C

    def one(r, c, ptr, idx, val):
        h = C.c_void_p()
        check(lib.ss_csr_wrap(ctx.h, r, c, G["nnz"], C.c_void_p(ptr.data_ptr()), C.c_void_p(idx.data_ptr()),
                              C.c_void_p(val.data_ptr()) if val is not None else None, C.byref(h)))
        return h
    return (one(G["ns"], G["nt"], G["y_ptr"], G["y_idx"], G["y_val"]),
            one(G["nt"], G["ns"], G["yt_ptr"], G["yt_idx"], G["yt_val"]))


def partial_products(G, s_end):
    """sum over the first s_end users of sum_{t' in Y[s]} sum_{s' in Y'[t']} ks[s']."""
    ks = (G["y_ptr"][1:] - G["y_ptr"][:-1]).to(torch.float64)
    per_item = torch.zeros(G["nt"], dtype=torch.float64, device=ks.device).index_add_(
        0, G["y_idx"].long(), ks.repeat_interleave((G["y_ptr"][1:] - G["y_ptr"][:-1]).long()))
    return float(per_item[G["y_idx"][: int(G["y_ptr"][s_end])].long()].sum().item())
