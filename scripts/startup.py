import time, sys
sys.path.insert(0, '.')
from vcfdist_b200 import capi
t = time.time(); e = capi.Engine(0); print("first vd_create %.0f ms" % ((time.time() - t) * 1e3))
t = time.time(); e2 = capi.Engine(0); print("second vd_create %.0f ms" % ((time.time() - t) * 1e3))
