import sys, ctypes
sys.path.insert(0,'/root/repo')
import lizard_b200 as lz
from tests import refs
from tests.test_gpu_frame import _mixed
BS=lz.BLOCK_SIZE
ref = lz.bind_frame_api(refs.ref_parity()); ours = lz.bind_frame_api(lz.lib())
level=21
data=_mixed(9*BS+12345, level)
p=lz.make_prefs(level,1,True,False,0)
fr=lz.frame_compress(ref,data,p)
r,back=lz.frame_decompress(ours,fr,len(data))
print("result",hex(r), ours.LizardF_getErrorName(r) if ours.LizardF_isError(r) else "", len(back), lz.lib().LizardB200_lastError())
# walk blocks
pos=7; i=0; units=[]; caps=[]
while True:
    w=int.from_bytes(fr[pos:pos+4],'little'); pos+=4
    if w==0: break
    sz=w&0x7fffffff
    print("block",i,"raw" if w>>31 else "comp",sz)
    if not (w>>31): units.append(fr[pos:pos+sz]); caps.append(BS)
    pos+=sz; i+=1
out=lz.decompress_batch(units,caps)
print([r for r,_ in out])
print([refs.ref_decompress(ref,u,BS)[0] for u in units])
