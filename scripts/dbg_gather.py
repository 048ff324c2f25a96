import sys, torch, threading
sys.path.insert(0, '.')
import accl_b200 as A
from accl_b200 import SUM
COUNT=5000
EAGER = dict(n_egr_rx_bufs=4, egr_rx_buf_size=16 << 10, max_egr_size=64 << 10, max_rndzv_size=1 << 30)
accls = A.cuda_world([0,0], heap_mb=64, max_ctas=4)
def body(r):
    torch.cuda.set_device(0)
    a = accls[r]
    with torch.cuda.stream(torch.cuda.Stream(0)):
        a.initialize(**EAGER)
        a.set_timeout(100000)  # ~3 s
        w = 2
        try:
            for root in range(w):
                for op in ("bcast","scatter","gather","reduce"):
                    print(f"r{r} root{root} {op} start", flush=True)
                    if op == "bcast":
                        b = a.create_buffer(COUNT); a.bcast(b, COUNT, root)
                    elif op == "scatter":
                        send, recv = a.create_buffer(COUNT * w), a.create_buffer(COUNT); a.scatter(send, recv, COUNT, root)
                    elif op == "gather":
                        out = a.create_buffer(COUNT * w); a.gather(recv, out, COUNT, root)
                    else:
                        s, d = a.create_buffer(COUNT), a.create_buffer(COUNT); a.reduce(s, d, COUNT, root, SUM)
                    print(f"r{r} root{root} {op} done", flush=True)
        except Exception as e:
            print(f"r{r} EXC {e}\n" + A._C.cuda_debug_state(a.impl), flush=True)
ts=[threading.Thread(target=body,args=(r,)) for r in range(2)]
[t.start() for t in ts]; [t.join() for t in ts]
for a in accls: print(A._C.cuda_debug_state(a.impl))
