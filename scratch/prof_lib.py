import sys, os
os.environ["VLG_B200_LIB"] = sys.argv[1]
sys.argv = [sys.argv[0]] + sys.argv[2:]
exec(open("scratch/prof_tc.py").read())
