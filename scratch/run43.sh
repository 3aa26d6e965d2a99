for lib in libmaddpg_b200.so libmaddpg_b200_u4.so libmaddpg_b200_u8.so; do
  MDP_LIB_NAME=$lib python bench.py --no-cpu-baseline --no-tensor-section --update-rounds 30 --e2e-steps 25 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$lib', 'value %.1fM' % (d['value']/1e6), 'upd %.0f (%.1f us/round)' % (d['critic_updates']['value'], 1e3*d['critic_updates']['ms_per_round']), 'grouped %.0f' % d['critic_updates']['grouped']['value'])"
done
