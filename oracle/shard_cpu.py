"""CPU stand-in for the device work of the sharded-table protocol (K9) --
TEST INFRASTRUCTURE ONLY.  Injected into ``ShardRouter`` by the gloo tests so the
exchange protocol (split sizes, all-to-all ordering, inverse permutation) runs
with world_size > 1 on CPU; also the oracle the CUDA partition kernel is checked
against."""
import torch


class CpuShardOps:
    def partition(self, ids_flat: torch.Tensor, world: int, rows_global: int):
        if ids_flat.numel() and (int(ids_flat.min()) < 0 or int(ids_flat.max()) >= rows_global):
            raise IndexError("embedding id out of range")
        owner = ids_flat % world
        order = torch.argsort(owner, stable=True)                 # stable: original order kept inside an owner group
        send_rows = ids_flat[order] // world
        inv = torch.empty_like(order)
        inv[order] = torch.arange(order.numel(), dtype=order.dtype)
        counts = torch.bincount(owner, minlength=world).to(torch.int32)
        return send_rows, order, inv, counts

    def take_rows(self, src, rows, width, stride, dtype_code, idx):
        return src[idx, :width].to(torch.float32).contiguous()
