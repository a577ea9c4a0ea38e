"""Layered boolean circuits on bootstrapped gates (BASELINE config #4; boolean.rs:9-53 per gate).

A circuit is a list of levels; a gate of level L reads two wires of level L-1 (level 0 = the encrypted
inputs).  Every gate is one `gate(op, ct0, ct1)` = PBS of 2*ct1 + ct0 (ct1 = LEFT input, boolean.rs:18).
Levels are evaluated one batched call at a time; with several ranks the gates of a level are sharded
across ranks (keys replicated) and the level's outputs are all-gathered so that the next level can read
any wire -- the only collective, once per level.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, List

import numpy as np

from . import sharding

AND, OR, XOR, NAND, NOR, XNOR = range(6)
_PLAIN = [lambda l, r: l & r, lambda l, r: l | r, lambda l, r: l ^ r,
          lambda l, r: 1 - (l & r), lambda l, r: 1 - (l | r), lambda l, r: 1 - (l ^ r)]


@dataclass
class Level:
    ops: np.ndarray   # uint8 [G]   gate opcodes
    left: np.ndarray  # int64 [G]   wire index (previous level) of the LEFT input  -> ct1 of boolean.rs
    right: np.ndarray  # int64 [G]  wire index (previous level) of the RIGHT input -> ct0 of boolean.rs


def random_layered_circuit(n_inputs: int, widths: List[int], seed: int, ops=(NAND, AND, XOR)) -> List[Level]:
    """gates uniform over `ops`, each reading two random wires of the previous level."""
    rng = np.random.default_rng(seed)
    levels, prev = [], n_inputs
    for w in widths:
        levels.append(Level(rng.choice(np.array(ops, dtype=np.uint8), w), rng.integers(0, prev, w), rng.integers(0, prev, w)))
        prev = w
    return levels


def evaluate_plain(levels: List[Level], inputs) -> np.ndarray:
    wires = np.asarray(inputs, dtype=np.int64)
    for lv in levels:
        l, r = wires[lv.left], wires[lv.right]
        wires = np.array([_PLAIN[op](int(a), int(b)) for op, a, b in zip(lv.ops, l, r)], dtype=np.int64)
    return wires


def evaluate_encrypted(levels: List[Level], wires, gate_fn: Callable, group=None):
    """wires: torch tensor [n_inputs, n+1] (every rank holds all wires of the current level).

    gate_fn(ops_uint8[g], ct0[g, n+1], ct1[g, n+1]) -> [g, n+1] evaluates a batch of gates on this rank.
    Returns the last level's wires (on every rank).
    """
    import torch
    rank, world = sharding.world_info(group)
    for lv in levels:
        G = len(lv.ops)
        lo, hi = sharding.shard_range(G, rank, world)
        if hi > lo:
            idx_r = torch.as_tensor(lv.right[lo:hi], device=wires.device)
            idx_l = torch.as_tensor(lv.left[lo:hi], device=wires.device)
            out = gate_fn(lv.ops[lo:hi], wires.index_select(0, idx_r).contiguous(), wires.index_select(0, idx_l).contiguous())
        else:
            out = wires.new_empty((0, wires.shape[1]))
        wires = sharding.all_gather_rows(out, G, group)
    return wires
