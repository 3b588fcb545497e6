import sys, time, ctypes as C, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import bench, ksw2_b200 as K, harness as H
import torch
mat=H.simple_mat(5,2,4)
def run(workload, n, mode, wp=0, steps=2):
    W=bench.WORKLOADS[workload]; q,qo,t,to=bench.gen(workload,n,0)
    dq=torch.from_numpy(q).cuda(); dt=torch.from_numpy(t).cuda()
    ctx=K.Context(0); ctx.set_mode(mode,wp)
    P=K.make_params(W['kind'],mat,**W['par']); L=K.lib()
    pl=L.ksw2b_plan_create(ctx.h,C.byref(P),n,qo.ctypes.data,to.ctypes.data)
    sp=C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for _ in range(2): L.ksw2b_plan_run(pl,dq.data_ptr(),dt.data_ptr(),None,sp)
    torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps): L.ksw2b_plan_run(pl,dq.data_ptr(),dt.data_ptr(),None,sp)
    e1.record(); torch.cuda.synchronize(); ms=e0.elapsed_time(e1)/steps
    cells=L.ksw2b_plan_cells(pl)
    print(f"{workload} n={n} mode={mode} wpanel={wp}: {ms:.1f} ms/step, {cells/ms/1e6:.1f} GCUPS", flush=True)
    L.ksw2b_plan_destroy(pl); ctx.close()
for n in (2000, 20000):
    run('c3', n, 1); run('c3', n, 2, 128); run('c3', n, 2, 256)
run('c3', 20000, 2, 64)
