import sys, os, time
sys.path.insert(0, os.getcwd())
import _pkg; rt=_pkg.load()
import numpy as np
def log(*a): print(*a, flush=True)
log("start", os.environ.get("RT_KERNEL"))
scene = rt.Scene.load_toml("assets/scenes/house.toml")
env = rt.EnvironmentMaps([rt.Environment.synthetic(256,128)])
log("inputs ready")
for (w,h,spp) in [(256,144,1),(256,144,2),(64,36,2),(300,200,3)]:
    st = rt.State(scene, env, w, h, max_bounces=8)
    log("state ready", w, h)
    t=time.time()
    try:
        st.render_samples(spp); log("launched"); st.sync()
        log(w,h,spp,"ok", round(time.time()-t,3), st.counters()["rays"])
    except Exception as e:
        log(w,h,spp,"ERR", e, round(time.time()-t,3))
    st.close()
