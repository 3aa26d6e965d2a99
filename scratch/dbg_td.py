import sys, numpy as np, torch
sys.path.insert(0, '.')
from tests.helpers import *
from tests.test_trainer_gpu import _build
name = sys.argv[1] if len(sys.argv) > 1 else "simple_spread"
case = trainer_case(name, seed=2)
oc = trainer_case(name, seed=2)
trainers, core = _build(case)
j = 0
idx = core.ring.index_tensor(case["idx"][j]); batch = core.ring.gather(idx); B = case["B"]
ut = torch.zeros((B, core.act_stride), device="cuda"); ut[:, :core.act_sum] = torch.from_numpy(case["u_target"][j]).cuda()
y, ta = core.td_target(j, batch, ut, want_target_act=True)
ta = ta.cpu().numpy()
# oracle pieces
fill_oracle_replay(oc)
otr = oc["trainers"]
n = oc["n"]; off = np.concatenate([[0], np.cumsum(oc["act_dims"])]).astype(int)
nobs = [otr[i].replay_buffer.sample_index(case["idx"][j])[3] for i in range(n)]
ota = []
for i in range(n):
    u = case["u_target"][j][:, off[i]:off[i+1]]
    otr[i].noise = lambda shape, u=u: u
    ota.append(otr[i].target_act(nobs[i]))
ota = np.concatenate(ota, 1)
d = np.abs(ta[:, :core.act_sum] - ota)
print("target act max diff", d.max(), "rows with diff>1e-4:", np.where(d.max(1) > 1e-4)[0])
bad = np.where(d.max(1) > 1e-4)[0]
for r in bad[:4]:
    c = d[r].argmax()
    i = np.searchsorted(off, c, side='right') - 1
    print("row", r, "agent", i, "gpu", ta[r, off[i]:off[i+1]], "ora", ota[r, off[i]:off[i+1]], "u", case["u_target"][j][r, off[i]:off[i+1]])
    lg = otr[i].target_p.forward(nobs[i])[0][r]
    print("   oracle logits", lg)
tq = otr[j].target_q_values(*(nobs + [ota[:, off[i]:off[i+1]] for i in range(n)]))
_, _, rew, _, done = otr[j].replay_buffer.sample_index(case["idx"][j])
yo = (rew + 0.95 * (1 - done) * tq).astype(np.float32)
dy = np.abs(y.cpu().numpy() - yo)
print("y max diff", dy.max(), np.where(dy > 1e-4)[0])
