# averages the per-SM lines a -DDMF_LINE_STATS build prints from dmf_counters() (tools/line_stats.py):  awk -f tools/line_stats_agg.awk log
/^sm /{n++; b+=$4; w+=$6; p+=$8; m+=$11; l+=$13; e+=$15}
END{printf "SMs %d  blocks/SM %.0f  warp-cycles/warp %.0f  prologue %.0f  march loop %.0f (line %.0f, exact %.0f)  epilogue %.0f\n", n, b/n, w/n, p/n, m/n, l/n, e/n, (w-p-m)/n}
