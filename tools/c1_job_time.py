import sys, os
sys.path.insert(0, "/root/repo")
import torch, bench_inputs, markovflow_b200 as mf
dev = torch.device("cuda:0")
ssm1, h1, y1, lr1 = bench_inputs.kalman_inputs_config3(1000, dev)
kf = mf.KalmanFilter(ssm1, mf.EmissionModel(h1), y1, lr1)
def job():
    post = kf.posterior_state_space_model()
    return kf.log_likelihood(), post.marginals
for _ in range(3): job()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): job()
e1.record(); torch.cuda.synchronize()
print("config-1 job (log-lik + posterior SSM + marginals): %.3f ms" % (e0.elapsed_time(e1) / 20))
