"""Synthetic scene description for benchmarks and parity tests: camera intrinsics, an object-model store
with the same attributes the inference path reads from datagen.ObjectModelStore
(/root/reference/epos_lib/datagen.py:24-154: dp_model['obj_ids'], frag_centers, frag_sizes), and "planted" network
outputs rendered analytically from known poses, because random-init heads carry no pose consensus."""
import numpy as np

SPHERE_RADIUS_MM = 100.0


def default_K():
    """YCB-V intrinsics (graph-cut-ransac/examples/example_pnp.ipynb)."""
    return np.array([[1066.778, 0.0, 312.9869], [0.0, 1067.487, 241.3109], [0.0, 0.0, 1.0]])


class ModelStore:
    """The subset of datagen.ObjectModelStore used by corresp.establish_many_to_many and infer.py."""

    def __init__(self, obj_ids, frag_centers, frag_sizes):
        self.dp_model = {'obj_ids': list(obj_ids)}
        self.frag_centers = frag_centers          # {obj_id: [F,3] f64}
        self.frag_sizes = frag_sizes              # {obj_id: [F] f64}
        self.num_frags = len(next(iter(frag_sizes.values())))

    def packed(self, num_objs):
        """Dense arrays indexed by obj_id-1 for the C ABI: centers [O,F,3], sizes [O,F] (zeros for unknown ids)."""
        F = self.num_frags
        c = np.zeros((num_objs, F, 3))
        s = np.ones((num_objs, F))
        for oid in self.dp_model['obj_ids']:
            if 1 <= oid <= num_objs:
                c[oid - 1] = self.frag_centers[oid]
                s[oid - 1] = self.frag_sizes[oid]
        return c, s


def _sphere_fragments(F, seed):
    """Farthest-point sampling of F centres among 10^4 uniform samples of a 100 mm sphere; fragment size = largest
    bounding-box side of the fragment's Voronoi cell, floored at 5 mm (datagen.py:112-122)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    p = rng.standard_normal((10000, 3))
    p *= SPHERE_RADIUS_MM / np.linalg.norm(p, axis=1, keepdims=True)
    idx = [0]
    d = np.linalg.norm(p - p[0], axis=1)
    for _ in range(1, F):
        k = int(np.argmax(d))
        idx.append(k)
        d = np.minimum(d, np.linalg.norm(p - p[k], axis=1))
    centers = p[idx]
    owner = np.argmin(((p[:, None, :] - centers[None]) ** 2).sum(-1), axis=1) if F * 10000 * 3 < 5e7 else \
        np.concatenate([np.argmin(((p[i:i + 500, None, :] - centers[None]) ** 2).sum(-1), axis=1)
                        for i in range(0, 10000, 500)])
    sizes = np.empty(F)
    for f in range(F):
        q = p[owner == f]
        sizes[f] = max(float((q.max(0) - q.min(0)).max()) if len(q) else 0.0, 5.0)
    return centers.astype(np.float64), sizes


def model_store(num_objs, num_frags, obj_ids=None):
    obj_ids = list(range(1, num_objs + 1)) if obj_ids is None else list(obj_ids)
    fc, fs = {}, {}
    for oid in obj_ids:
        fc[oid], fs[oid] = _sphere_fragments(num_frags, seed=oid)
    return ModelStore(obj_ids, fc, fs)


def rodrigues(rv):
    th = float(np.linalg.norm(rv))
    if th < 1e-12:
        return np.eye(3)
    k = np.asarray(rv, np.float64) / th
    Kx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + np.sin(th) * Kx + (1 - np.cos(th)) * (Kx @ Kx)


def planted_maps(B, num_objs, num_frags, store, K, seed=0, objs_per_image=None, outlier_frac=0.3, loc_noise=0.02,
                 h=120, w=160, output_scale=0.25, instances_per_object=1):
    """Model outputs with known poses.  Returns (obj_conf [B,h,w,O+1], frag_conf [B,h,w,O,F], frag_loc [B,h,w,O,F,3],
    gt) with gt[b] = {obj_id: (R, t)}.  Each visible object is the 100 mm sphere of `store` at t_z in [500, 1200] mm;
    inside its silhouette obj_conf = 0.9, frag_conf = 0.8 on the true fragment, frag_loc = (X - centre)/size + noise;
    `outlier_frac` of the silhouette pixels get a random fragment and random local coordinates instead.
    instances_per_object > 1 plants that many spheres per visible object (later ones overwrite earlier ones where they
    overlap); gt[b][obj_id] is then a LIST of (R, t)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    O, F = num_objs, num_frags
    obj_conf = np.zeros((B, h, w, O + 1), np.float32)
    obj_conf[..., 0] = 1.0
    frag_conf = np.full((B, h, w, O, F), 1.0 / F, np.float32)
    frag_loc = (rng.standard_normal((B, h, w, O, F, 3)) * 0.01).astype(np.float32)
    Kinv = np.linalg.inv(K)
    ys, xs = np.mgrid[0:h, 0:w]
    u = (xs + 0.5) / output_scale
    v = (ys + 0.5) / output_scale
    rays = np.stack([u, v, np.ones_like(u)], -1) @ Kinv.T                      # [h,w,3]
    gt = []
    ids = store.dp_model['obj_ids']
    for b in range(B):
        g = {}
        n_vis = len(ids) if objs_per_image is None else min(objs_per_image, len(ids))
        vis = list(rng.choice(ids, size=n_vis, replace=False)) if n_vis < len(ids) else list(ids)
        for oid in [o for o in vis for _ in range(instances_per_object)]:
            tz = rng.uniform(500.0, 1200.0)
            cu, cv = rng.uniform(0.2, 0.8) * w / output_scale, rng.uniform(0.2, 0.8) * h / output_scale
            t = tz * (Kinv @ np.array([cu, cv, 1.0]))
            rv = rng.standard_normal(3)
            rv *= rng.uniform(0.2, 2.8) / np.linalg.norm(rv)
            R = rodrigues(rv)
            # ray-sphere intersection |s d - t|^2 = r^2 (nearest root)
            dd = (rays * rays).sum(-1)
            dt = rays @ t
            disc = dt * dt - dd * (t @ t - SPHERE_RADIUS_MM ** 2)
            hit = disc > 0
            s = (dt - np.sqrt(np.where(hit, disc, 0.0))) / dd
            Xc = rays * s[..., None]
            Xo = (Xc - t) @ R                                                    # R^T (Xc - t)
            cen, siz = store.frag_centers[oid], store.frag_sizes[oid]
            yy, xx = np.nonzero(hit)
            if len(yy) == 0:
                continue
            P = Xo[yy, xx]
            f = np.argmin(((P[:, None, :] - cen[None]) ** 2).sum(-1), axis=1)
            loc = (P - cen[f]) / siz[f][:, None] + rng.standard_normal(P.shape) * loc_noise
            out = rng.random(len(yy)) < outlier_frac
            f = np.where(out, rng.integers(0, F, len(yy)), f)
            loc = np.where(out[:, None], rng.uniform(-0.5, 0.5, P.shape), loc)
            obj_conf[b, yy, xx, :] = 0.1 / O
            obj_conf[b, yy, xx, oid] = 0.9
            frag_conf[b, yy, xx, oid - 1, :] = 0.2 / (F - 1) if F > 1 else 1.0
            frag_conf[b, yy, xx, oid - 1, f] = 0.8 if F > 1 else 1.0
            frag_loc[b, yy, xx, oid - 1, f, :] = loc.astype(np.float32)
            if instances_per_object == 1:
                g[int(oid)] = (R, t)
            else:
                g.setdefault(int(oid), []).append((R, t))
        gt.append(g)
    return obj_conf, frag_conf, frag_loc, gt
