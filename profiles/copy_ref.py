import torch
for shape in [(256,256,256,8),(256,128,128,16),(256,64,64,32)]:
    x = [torch.randn(*shape, device="cuda") for _ in range(2)]; y = [torch.empty_like(x[0]) for _ in range(2)]
    for i in range(3): y[i%2].copy_(x[i%2])
    e0,e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for i in range(10): y[i%2].copy_(x[i%2])
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1)/10*1e3
    print(shape, f"copy {us:.1f} us, {2*x[0].numel()*4/us/1e3:.0f} GB/s")
