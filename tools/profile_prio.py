"""A few rounds of the device prioritized replay at the reference's capacity (for an ncu launch list).
   python tools/profile_prio.py"""
import sys, torch
sys.path.insert(0, '.')
from maddpg_b200 import DevicePrioritizedReplayMemory
cap, E, B = 1000000, 4096, 1024
mem = DevicePrioritizedReplayMemory(cap, numpy_io=False, strict=False)
z = lambda *s: torch.zeros(s, device="cuda")
for _ in range(25):
    mem.add(z(E, 18), z(E, 5), z(E), z(E, 18), torch.zeros(E, dtype=torch.uint8, device="cuda"))
u = torch.rand(B, dtype=torch.float64, device="cuda")
err = torch.rand(B, dtype=torch.float64, device="cuda")
for _ in range(3):
    tidx, _, _ = mem.sample(B, uniforms=u)
    mem.batch_update(tidx, err)
    mem.add(z(E, 18), z(E, 5), z(E), z(E, 18), torch.zeros(E, dtype=torch.uint8, device="cuda"))
torch.cuda.synchronize()
print("ok")
