"""Sharded contraction Pi against the single-GPU one, random and kronecker-structured index sets: python tools/shard_check.py NGPU"""
import sys, time
sys.path.insert(0, ".")
import numpy as np
import tci_b200 as T
import bench as B
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
cn, c1 = T.Context(devices=list(range(n))), T.Context(0)
A, Bc = B.mpo_cores(5, nsites=16, D=64), B.mpo_cores(6, nsites=16, D=64)
fn = T.Contraction(T.TensorTrain(A), T.TensorTrain(Bc), ctx=cn)
f1 = T.Contraction(T.TensorTrain(A), T.TensorTrain(Bc), ctx=c1)
rng = np.random.default_rng(1)
base_I = np.stack([rng.integers(1, 5, 128) for _ in range(7)], axis=1).astype(np.int64)
base_J = np.stack([rng.integers(1, 5, 128) for _ in range(7)], axis=1).astype(np.int64)
sets = {"random": (np.stack([rng.integers(1, 5, 500) for _ in range(8)], axis=1).astype(np.int64),
                   np.stack([rng.integers(1, 5, 300) for _ in range(8)], axis=1).astype(np.int64)),
        "kronecker": (T.kronecker_left(np.unique(base_I, axis=0), 4), T.kronecker_right(4, np.unique(base_J, axis=0)))}
import os
os.environ["TCI_SHARD_FORCE"] = str(n)
for name, (I, J) in sets.items():
    a, b = fn(I, J, 0), f1(I, J, 0)
    da, ma = fn.batchevaluate_device(I, J, 0)
    db, mb = f1.batchevaluate_device(I, J, 0)
    print(name, I.shape, J.shape, "max rel dev", float(np.max(np.abs(a - b)) / np.max(np.abs(b))), "maxabs equal", ma == mb,
          "member launches", [cn.member_launches(k) for k in range(n)], file=sys.stderr)
