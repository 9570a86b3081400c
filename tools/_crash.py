import faulthandler; faulthandler.enable()
import os, sys; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qboot_b200 as qb, sys
t=int(sys.argv[1]) if len(sys.argv)>1 else 8
print(qb.build_pmp(600,60,9,'3',0,0,True,numax=6,maxspin=12,threads=t,outdir='/tmp/x_ours'))
