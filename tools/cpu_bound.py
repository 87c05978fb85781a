import sys, time, torch, numpy as np
sys.path.insert(0, '/root/repo')
import tf_image_compression_b200 as T
MEAN = np.array([118.3, 113.9, 102.6], np.float32); STD = np.array([61.7, 59.2, 63.8], np.float32)
codec = T.Codec("model_0", quan_scale=2, mean=MEAN, std=STD, compute="tensor")
codec.use_torch_stream(); codec.set_chunk_patches(12288)
B=64; img = torch.randint(0,256,(B,1536,2048,3),dtype=torch.uint8,device='cuda')
sym = torch.empty((B,192,8,8,64),dtype=torch.uint8,device='cuda'); rec=torch.empty_like(img)
for _ in range(3): codec.encode_images(img,128,out=sym); codec.decode_images(sym,1536,2048,128,out=rec)
torch.cuda.synchronize()
l0=codec.launch_count
t0=time.perf_counter(); codec.encode_images(img,128,out=sym); t1=time.perf_counter(); torch.cuda.synchronize(); t2=time.perf_counter()
codec.decode_images(sym,1536,2048,128,out=rec); t3=time.perf_counter(); torch.cuda.synchronize(); t4=time.perf_counter()
print(f"encode: host issue {1e3*(t1-t0):.2f} ms, total {1e3*(t2-t0):.2f} ms; decode: host issue {1e3*(t3-t2):.2f} ms, total {1e3*(t4-t2):.2f} ms; launches {codec.launch_count-l0}")
