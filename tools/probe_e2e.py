"""e2e (host buffers, copies inside the step) vs device-resident step time, with and without the copy-stream staging."""
import sys, time
import torch
sys.path.insert(0, ".")
import bench
from clipa_b200 import open_clip
from clipa_b200.training import TrainStep
name, B = (sys.argv[1], int(sys.argv[2])) if len(sys.argv) > 2 else ("vitb16_i64_t16_gb16k", 2048)
wl = bench.WORKLOADS[name]
dev = torch.device("cuda:0")
torch.manual_seed(0)
model, _, _ = open_clip.create_model_and_transforms(wl["model"], precision="amp_bf16", device=dev, force_image_size=wl["image"],
                                                    pos_embed=wl["pos"], output_dict=True)
model.train()
ts = TrainStep(model, micro_batch=B)
g = torch.Generator().manual_seed(1)
h_img = torch.randint(0, 256, (B, 3, wl["image"], wl["image"]), generator=g, dtype=torch.uint8).pin_memory()
h_txt = torch.randint(1, model.vocab_size - 1, (B, model.context_length), generator=g); h_txt[:, -1] = model.vocab_size - 1
h_txt = h_txt.pin_memory()
d_img, d_txt = ts.preprocess(h_img), h_txt.to(dev)
def run(fn, n=6):
    for _ in range(3): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
print(f"PERF e2e {name} B={B}: device-resident {run(lambda: ts.step(d_img, d_txt)):.2f} ms", end="")
for mode in (True, False, True):
    ts.stage_on_copy_stream = mode
    print(f" | host buffers + item(), staging={mode}: {run(lambda: ts.step(h_img, h_txt).item()):.2f} ms", end="")
print()
t0 = time.perf_counter()
for _ in range(3): ts.step(d_img, d_txt)
host = (time.perf_counter() - t0) / 3 * 1e3
torch.cuda.synchronize()
print(f"PERF host enqueue time per step (no sync): {host:.2f} ms")
