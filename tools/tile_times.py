import sys; import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from rusterizer_b200 import scenes
from rusterizer_b200.render import Renderer
which = os.environ.get('RZ_SCENE', 'c2')
sc = {'c2': scenes.sphere_scene, 'c4ii': scenes.fullscreen_quad_scene, 'c4i': lambda: scenes.sphere_scene(width=8192, height=8192), 'c3': scenes.near_clip_scene, 'c1': scenes.default_scene, 'overdraw': scenes.overdraw_scene}[which]()
r = Renderer(sc.width, sc.height); r.uniforms().bind_texture(0, sc.texture)
dm = [r.upload(d.mesh) for d in sc.draws]
r.debug_capture(True)
for i in range(3):
    scenes.render_scene(r, sc, dm); r.framebuffer_device()
print(r.timings())
t = r.tile_times(); t = t[t[:,2] > 0]
n = (t[:,0] >> np.uint64(32)).astype(int); dur = (t[:,2]-t[:,1]).astype(float)/1e3
t0 = t[:,1].min(); print("tiles", len(t), "span us", (t[:,2].max()-t0)/1e3)
print("dur us: mean %.1f p50 %.1f p90 %.1f p99 %.1f max %.1f" % (dur.mean(), np.percentile(dur,50), np.percentile(dur,90), np.percentile(dur,99), dur.max()))
o = np.argsort(-dur)[:12]
for k in o: print("  tile", int(t[k,0] & np.uint64(0xffffffff)), "n", n[k], "dur %.1f start %.1f sm %d" % (dur[k], (t[k,1]-t0)/1e3, t[k,3]))
for lo,hi in [(1,32),(32,64),(64,128),(128,256),(256,512),(512,4096)]:
    m = (n>=lo)&(n<hi)
    if m.any(): print(f"n in [{lo},{hi}): tiles {m.sum():5d} mean dur {dur[m].mean():6.1f} us  total {dur[m].sum()/1e3:7.2f} ms")
ph = np.stack([((t[:,4] >> np.uint64(16*k)) & np.uint64(0xFFFF)).astype(float)*16/1e3 for k in range(4)] + [t[:,5].astype(float)/1e3], 1)
m = (n >= 64) & (n < 128) if which == 'c2' else (n >= 1)
names = ["A0 done", "A1 done", "A2 done", "B done", "C done"]
print("first-chunk phase completion (us after tile start), tiles with 64<=n<128:")
for k in range(5): print(f"   {names[k]:8s} mean {ph[m,k].mean():6.2f}  p50 {np.percentile(ph[m,k],50):6.2f}")
print("   tile end  mean %6.2f" % dur[m].mean())
print("sum of durations ms", dur.sum()/1e3, " / (444 slots) = us", dur.sum()/444)
# per SM busy time
sm = t[:,3].astype(int)
bus = np.bincount(sm, weights=dur)
print("per-SM sum dur: min %.0f mean %.0f max %.0f (x1/3 if 3 CTAs overlap)" % (bus[bus>0].min(), bus[bus>0].mean(), bus.max()))
