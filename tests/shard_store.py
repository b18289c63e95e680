"""TEST INFRASTRUCTURE — the CPU oracle behind the store-class protocol ShardedCube talks to, so that the
multi-rank host logic (row sharding, exchange bookkeeping, ordered combine) runs over gloo on machines
without a GPU.  The product (olap_in_memory_b200/sharded.py) knows nothing about it: it only calls the
protocol that GpuStore implements on the device — the batched `*_lowered` transforms, `recv_like`,
`exchange_planes` / `exchange_done`, `average_of`, `evaluate`, `DEVICE`, `PEER_MEMORY`."""
import numpy as np
import torch

from oracle.store_oracle import OracleStore


def _identity_maps(maps, lens):
    return [np.arange(n, dtype=np.int32) if m is None else m for m, n in zip(maps, lens)]


class OracleShardStore(OracleStore):
    DEVICE = "cpu"
    PEER_MEMORY = False  # no CUDA IPC here: rollups of a sharded dimension take the all-to-all exchange

    # ---- data boundary as the device store spells it
    def set_data_f32(self, values):
        self.data = [float(v) for v in np.asarray(values, dtype=np.float32)]

    def data_f32(self):
        """The protocol's bulk read.  The oracle keeps JS doubles: they are handed out unrounded, so that the
        gloo tests can hold the sharded host logic to 1e-12 against ONE oracle cube."""
        return np.asarray(self.data, dtype=np.float64)

    def _as_mine(self, other):
        out = OracleShardStore(other.size, other._type, other._defaultValue)
        out._dataMap = other._dataMap
        return out

    # ---- batched static forms (one call for all measures of a cube)
    @staticmethod
    def drillUp_lowered(stores, old_len, new_len, maps, methods):
        out = []
        for s, method in zip(stores, methods):
            full = [np.zeros(o, dtype=np.int32) if (m is None and o != n) else m for m, o, n in zip(maps, old_len, new_len)]
            full = _identity_maps(full, old_len)
            if method == "__count":  # number of set children: roll an indicator up
                ind = OracleShardStore(s.size, s._type, s._defaultValue)
                present = set(s._dataMap.keys())
                ind.data = [1.0 if i in present else s._defaultValue for i in range(s.size)]
                out.append(s._as_mine(OracleStore.drillUp_lowered(ind, old_len, new_len, full, "sum")))
            else:
                out.append(s._as_mine(OracleStore.drillUp_lowered(s, old_len, new_len, full, method)))
        return out

    @staticmethod
    def drillDown_lowered(stores, old_len, new_len, maps, methods, distributions=None):
        return [s._as_mine(OracleStore.drillDown_lowered(s, old_len, new_len, maps, meth)) for s, meth in zip(stores, methods)]

    @staticmethod
    def dice_lowered(stores, old_len, keep):
        return [s._as_mine(OracleStore.dice_lowered(s, old_len, keep)) for s in stores]

    @staticmethod
    def reorder_lowered(stores, old_len, new_to_old):
        return [s._as_mine(OracleStore.reorder_lowered(s, old_len, new_to_old)) for s in stores]

    @staticmethod
    def evaluate_to_store(expression, cell_names, stores, totals, type="float32", defaultValue=0):
        size = stores[0].size
        out = OracleShardStore(size, type, defaultValue)
        out.data = OracleStore.evaluate(expression, cell_names, stores, totals, size)
        return out

    # ---- exchange and combine
    @classmethod
    def recv_like(cls, store, size):
        return cls(size, store._type, store._defaultValue)

    @staticmethod
    def exchange_planes(store):
        return [torch.tensor(store.data, dtype=torch.float64)]

    @staticmethod
    def exchange_done(store, planes):
        store.data = planes[0].tolist()

    @classmethod
    def average_of(cls, sums, counts):
        out = cls(sums.size, sums._type, sums._defaultValue)
        out.data = [sv / cv if (cv == cv and cv != 0) else sums._defaultValue for sv, cv in zip(sums.data, counts.data)]
        return out
