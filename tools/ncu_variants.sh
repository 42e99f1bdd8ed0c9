M="--metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sectors_srcunit_tex_op_read.sum --clock-control none -k regex:conv_tc_kernel -s 2 -c 1"
for v in "" "LDB200_NO_TMA_EPI=1" "LDB200_NO_L2_PREFETCH=1" "LDB200_NO_TMA_EPI=1 LDB200_NO_L2_PREFETCH=1"; do echo "== $v res"; env $v timeout 60 ncu $M python tools/one_conv.py 16 200 336 64 256 1 1 0 res 2>&1 | grep -E "dram|duration|lts"; done
echo "== plain"; timeout 60 ncu $M python tools/one_conv.py 16 200 336 64 256 1 1 0 2>&1 | grep -E "dram|duration|lts"
