import re,subprocess,sys
log=open(sys.argv[1] if len(sys.argv)>1 else 'vkradixsort_b200/lib/ptxas.log').read()
ents=re.findall(r"Compiling entry function '(\S+)' for 'sm_100a'\n.*?\n.*?(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n.*?Used (\d+) registers(.*)", log)
for name,stack,ss,sl,regs,rest in ents:
    dem=subprocess.run(['c++filt',name],capture_output=True,text=True).stdout.strip()
    dem=re.sub(r'\(.*','',dem)
    print(f"{regs:>4} regs stack={stack:>4} spill={ss}/{sl}  {dem[5:110]}")
