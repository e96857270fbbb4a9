/^sm /{n++; w+=$7; p+=$9; m+=$12; l+=$14; e+=$16}
END{printf "SMs %d  warp-cycles/warp %.0f  prologue %.0f  march loop %.0f (line %.0f exact %.0f)  rest(epilogue) %.0f\n", n, w/n, p/n, m/n, l/n, e/n, (w-p-m)/n}
