#!/bin/bash
# regenerates tests/golden/ocr/ from the reference checkout (run in the build container, /root/reference is absent on the GPU box)
set -e
here=$(cd "$(dirname "$0")" && pwd)
mkdir -p "$here/ocr"
cp /root/reference/misc/textline.bin.png /root/reference/misc/textline.gt.txt "$here/ocr/"
