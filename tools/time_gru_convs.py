"""cuDNN TF32 convolution time for the candidate ConvGRU formulations (channels-last, B8 48x156)."""
import statistics, torch, torch.nn.functional as F
torch.backends.cudnn.allow_tf32 = True
torch.backends.cudnn.benchmark = True
cl = torch.channels_last
def t(cin, cout, k, reps=20):
    x = torch.randn(8, cin, 48, 156, device="cuda").contiguous(memory_format=cl)
    w = torch.randn(cout, cin, *k, device="cuda").contiguous(memory_format=cl)
    pad = (k[0] // 2, k[1] // 2)
    for _ in range(5): F.conv2d(x, w, None, padding=pad)
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); F.conv2d(x, w, None, padding=pad); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return statistics.median(ts)
for name, cin, couts in (("3x input  [hi;lo;hi] Cin=1152", 1152, (256, 128)), ("2x input  [hi;hi]    Cin=768 ", 768, (256, 128)),
                         ("2x output [w_hi;w_lo] Cin=384 ", 384, (512, 256)), ("plain TF32           Cin=384 ", 384, (256, 128)),
                         ("2x input padded      Cin=1024", 1024, (256, 128))):
    for k in ((1, 5), (5, 1)):
        print(f"{name} k={k}: zr {t(cin, couts[0], k):7.1f} us   q {t(cin, couts[1], k):7.1f} us", flush=True)
