"""oracle/bin2inc.py -- emit `unsigned char NAME[] = {...}; unsigned int NAME_len = N;` for (file, NAME) pairs.
Used by oracle/Makefile to embed the reference's matrix data files into oracle/_ref/ (git-ignored)."""
import sys
args = sys.argv[1:]
for path, name in zip(args[0::2], args[1::2]):
    data = open(path, "rb").read()
    print("static const unsigned char %s[] = {%s};" % (name, ",".join(str(b) for b in data)))
    print("static const unsigned int %s_len = %d;" % (name, len(data)))
