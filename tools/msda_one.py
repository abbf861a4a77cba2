import os, sys
sys.path.insert(0, '/root/repo')
sys.argv = ['x']
import tools.msda_perf as mp
mp.run(2, 'encoder')
